// vk_internal.h -- device-side scene layout and the launcher interface between vk_api.cu (context,
// upload, C ABI) and vk_kernels.cu (device code, compiled twice: namespace vkfast with FMA
// contraction, namespace vkstrict with -fmad=false and the reference's op sequence).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/vecchio_gpu.h"

// Device references add two things to vk_ref: bit 27 marks the SECOND visit of a single-object BVH
// leaf that holds a ConstantMedium (src/accel.rs:102-107), and type 15 is the
// traversal's "leave instance" sentinel.
#define VKD_DUP 0x08000000u
#define VKD_INDEX(r) ((r)&0x07FFFFFFu)
#define VKD_TYPE(r) ((r) >> 28)
#define VKD_T_EXIT 15u
#define VKD_STACK 64

// Scene arrays in HBM.  Every record is a multiple of 16 B and fetched with 128-bit loads
// through the read-only path; one array per primitive kind (struct-of-arrays by type).
struct DScene {
    // Traversal layout ("wide" node, 64 B = 4 x float4, one 128-byte-line half): the boxes of BOTH
    // children live in the parent, so one fetch decides both descents and the nearer child is
    // visited first.  {lmin.xyz, left}, {lmax.xyz, right}, {rmin.xyz, -}, {rmax.xyz, -}.  A box is
    // only meaningful for a child that is itself a node: the reference tests no box for a primitive
    // child (src/accel.rs:64-65 calls its hit() directly).
    const float4* wnodes;
    const float4* nodes;    // reference layout, 2 x float4 per node: {min.xyz, left}, {max.xyz, right}; used for
                            // the box of a BVH root (world root, instanced sub-BVH root)
    const float4* spheres;  // {center.xyz, radius}
    const uint32_t* sphere_mat;
    const float4* mspheres; // 3 x float4: {c0, r}, {c1, time0}, {time1, mat, -, -}
    const float4* rects;    // 2 x float4: {c0,c1,d0,d1}, {k, axes, mat, -}
    const float4* boxes;    // 2 x float4: {min.xyz, mat}, {max.xyz, -}
    const float4* xforms;   // 2 x float4: {kind, child, -, -}, {a, b, c, -}
    const float4* media;    // {boundary, neg_inv_density, mat, -}
    const uint32_t* lights;
    const uint4* materials; // {type, tex, param, aux | needs_uv << 31}
    const uint4* textures;
    const uint8_t* texels;
    const float4* perlin_vec;   // 256 x float4 per Perlin
    const uint8_t* perlin_perm; // 768 B per Perlin (x, y, z)
    uint32_t root;
    uint32_t n_lights;
    uint32_t has_media; // scene holds a ConstantMedium: selects the kernel instantiation with the medium code
};
#define VKD_MAT_NEEDS_UV 0x80000000u

// ------------------------------------------------------------------------------------------------
// Flat traversal program.  A scene with few primitives (Cornell box: 13 rect sides + 1 sphere) gains
// nothing from its BVH on a GPU: the lanes of a warp sit at different nodes and execute each
// other's branches.  For such scenes vk_scene_upload unrolls the reference's traversal order
// (depth first, left then right, a Boxy as its six sides, a wrapper chain as push/pop of the ray
// frame) into a short straight-line program.  Every lane then tests the SAME primitive at the same
// time: no stack, no divergence, warp-uniform operand fetches from the constant bank (the program
// travels as a __grid_constant__ kernel parameter).  Closest hit is order independent, so this is
// the same function as BVHNode::hit; the BVH path remains for everything larger.
// ------------------------------------------------------------------------------------------------
#define VKD_FLAT_MAX 96
enum {
    VKF_RECT_XY = 1, VKF_RECT_XZ, VKF_RECT_YZ, VKF_SPHERE, VKF_MSPHERE, VKF_MEDIUM,
    VKF_PUSH_TRANSLATE, VKF_PUSH_ROTX, VKF_PUSH_ROTY, VKF_PUSH_ROTZ, VKF_POP
};
#define VKF_STRICT 0x100u // a side of a Boxy: list semantics, must be strictly closer (src/hittable.rs:386)
struct FlatEntry {
    float4 a;      // rect: c0,c1,d0,d1 | sphere: center,r | msphere: c0,r | translate: offset | rotate: sin,cos
    float4 b;      // msphere: c1,time0
    float k;       // rect: plane | msphere: time1
    uint32_t kind; // VKF_*
    uint32_t ref;  // leaf record (hit id) | push: the chain's outermost wrapper (instance id)
    uint32_t aux;  // face (0..5) | VKF_STRICT
};
struct FlatProgram {
    uint32_t n; // 0 = no program: use the BVH
    uint32_t _pad[3];
    FlatEntry e[VKD_FLAT_MAX];
};

struct DCamera {
    float3 origin, lower_left_corner, horizontal, vertical, u, v;
    float lens_radius, time0, time1;
};

struct RenderArgs {
    uint32_t width, height;
    uint32_t spp_begin, spp_count; // global sample range of this call
    uint32_t max_depth;
    uint32_t seed_lo, seed_hi;
    float3 background;
    // work decomposition (vk_api.cu decides): warp items = 8x4 pixel tiles x sample chunks; inside
    // an item the lanes share (pixel, block of unit_spp samples) units; one partial plane per block
    uint32_t tiles_x, tiles_y, n_chunks, chunk_spp, unit_spp, n_planes;
};

// counters[0] = rays, [1] = dropped samples, [2] = work-queue head, [3] = node visits, [4] = primitive tests
struct RenderBuffers {
    float* partial_sum;   // n_planes x W*H*3 (== d_sum when n_planes == 1)
    float* partial_sumsq; // same, nullable
    unsigned long long* counters;
};

// ------------------------------------------------------------------------------------------------
// Wavefront variant: a pool of path slots in device memory (L2-resident at the default pool size),
// advanced one segment per iteration by separate kernels:
//   generate  k_wf_generate  every slot takes a (pixel, sample block) unit and its first camera ray
//   extend    k_wf_extend    world.hit() for every live slot; the hit is classified by the material's
//                            shading class and the slot index appended to that class's queue
//                            (warp ballot + one atomic per warp and class)
//   shade     k_wf_shade     walks the queues class by class, so a warp shades one kind of material:
//                            scatter + PDF sampling, or sample end -> accumulate -> next sample / next
//                            unit (regeneration in place: the pool stays full until units run out)
// A slot sums the samples of its unit in order and stores the unit's partial sum to the plane of
// its sample block, exactly like a lane of the megakernel: both variants produce the same image
// for a seed.  float4 / uint4 arrays: every access is one fully coalesced 16 B per lane.
// ------------------------------------------------------------------------------------------------
enum { VKW_TERMINATE = 0, VKW_DIELECTRIC = 1, VKW_METAL = 2, VKW_DIFFUSE = 3, VKW_CLASSES = 4 };
struct WfState {
    float4* ray_o;  // origin.xyz, time
    float4* ray_d;  // direction.xyz, bits: depth of the segment to trace (0 = slot idle)
    float4* beta;   // path weight.xyz, bits: global index of the current sample
    uint4* unit;    // pixel, s_end (one past the unit's last sample), plane, -
    float4* sum;    // partial sum of the unit's finished samples
    float4* sumsq;  // same for squares; nullptr when not wanted
    uint4* hit;     // t bits, primitive, instance, face
    uint32_t* queue;    // VKW_CLASSES x n_slots slot indices
    uint32_t* qcount;   // 2 sets x VKW_CLASSES counters (set = iteration parity)
    unsigned long long* unit_head; // next unit to hand out
    uint32_t n_slots;
    uint32_t n_pixels;
    unsigned long long n_units;
};

#define VK_DECLARE_LAUNCHERS(NS)                                                                                       \
    namespace NS {                                                                                                     \
    cudaError_t launch_megakernel(const DScene& sc, const FlatProgram* flat, const DCamera& cam, const RenderArgs& a,  \
                                  const RenderBuffers& b, int grid, cudaStream_t st);                                  \
    cudaError_t launch_intersect(const DScene& sc, const FlatProgram* flat, const vk_ray* rays, size_t n,              \
                                 const float* medium_xi, vk_hit* out, cudaStream_t st);                                \
    cudaError_t megakernel_occupancy(bool flat, bool media, int* blocks_per_sm, int* block_threads);                               \
    cudaError_t launch_philox_kat(const uint32_t* ctr_key6, uint32_t* out4, cudaStream_t st);                          \
    cudaError_t launch_wf_generate(const DCamera& cam, const RenderArgs& a, const WfState& w, cudaStream_t st);        \
    cudaError_t launch_wf_extend(const DScene& sc, const FlatProgram* flat, const RenderArgs& a, const WfState& w,     \
                                 const RenderBuffers& b, uint32_t set, cudaStream_t st);                               \
    cudaError_t launch_wf_shade(const DScene& sc, const DCamera& cam, const RenderArgs& a, const WfState& w,           \
                                const RenderBuffers& b, uint32_t set, cudaStream_t st);                                \
    }
VK_DECLARE_LAUNCHERS(vkfast)
VK_DECLARE_LAUNCHERS(vkstrict)
