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
    const float4* nodes;    // 2 x float4 per node: {min.xyz, left}, {max.xyz, right}
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
};
#define VKD_MAT_NEEDS_UV 0x80000000u

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
    // work decomposition (vk_api.cu decides): warp tiles of 32 pixels x sample chunks
    uint32_t tiles_x, tiles_y, n_chunks, chunk_spp;
};

// counters[0] = rays, [1] = dropped samples, [2] = work-queue head
struct RenderBuffers {
    float* partial_sum;   // n_chunks x W*H*3 (== d_sum when n_chunks == 1)
    float* partial_sumsq; // same, nullable
    unsigned long long* counters;
};

#define VK_DECLARE_LAUNCHERS(NS)                                                                                       \
    namespace NS {                                                                                                     \
    cudaError_t launch_megakernel(const DScene& sc, const DCamera& cam, const RenderArgs& a, const RenderBuffers& b,  \
                                  int grid, cudaStream_t st);                                                          \
    cudaError_t launch_intersect(const DScene& sc, const vk_ray* rays, size_t n, const float* medium_xi, vk_hit* out,  \
                                 cudaStream_t st);                                                                     \
    cudaError_t megakernel_occupancy(int* blocks_per_sm, int* block_threads);                                          \
    cudaError_t launch_philox_kat(const uint32_t* ctr_key6, uint32_t* out4, cudaStream_t st);                          \
    }
VK_DECLARE_LAUNCHERS(vkfast)
VK_DECLARE_LAUNCHERS(vkstrict)
