// vk_api.cu -- the C ABI of include/vecchio_gpu.h: context, scene validation + upload (with the
// GPU-side re-layout), render orchestration, the parity hook and the roofline microbenchmarks.
// No CPU fallback anywhere: every entry point either runs CUDA kernels or returns an error.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <string>
#include <vector>

#include "vk_internal.h"

struct vk_ctx {
    int device = 0;
    int sm_count = 0;
    int clock_khz = 0;
    char name[128] = {0};
    cudaStream_t stream = nullptr;     // stream all work is enqueued on
    cudaStream_t own_stream = nullptr; // the one created by vk_create
    uint64_t launches = 0;             // kernels launched since the last stats read
    uint64_t paths = 0;                // samples started since the last stats read
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr;
    std::string err;
    std::vector<void*> scene_allocs;
    DScene scene{};
    FlatProgram flat{}; // flat.n == 0: scene too large, BVH traversal
    bool has_scene = false;
    uint32_t n_nodes = 0; // BVH nodes of the uploaded scene
    bool has_specdiffuse = false;
    bool simple_scene = false; // only what the VK_SIMPLE build of the staged kernel keeps (see vk_device.cuh)
    unsigned long long* debug = nullptr;    // per-CTA diagnostics of the staged kernel (VK_DEBUG_CTAS x 4 words)
    unsigned long long* counters = nullptr; // [0] rays [1] dropped [2] work head
    float* partial = nullptr;               // chunk partial sums (sum | sumsq)
    size_t partial_floats = 0;
    float* frame = nullptr; // vk_render staging: sum | sumsq | rgb
    size_t frame_floats = 0;
    float* pinned = nullptr; // pinned host staging for the D2H of vk_render
    size_t pinned_floats = 0;
    // wavefront variant: the slot pool (allocated on first use) and a pinned word block for the
    // host's termination checks
    WfState wf{};
    void* wf_block = nullptr;
    bool wf_has_sumsq = false;
    uint32_t* wf_host_counts = nullptr;
};

static thread_local std::string g_create_err;

static int fail(vk_ctx* c, int code, const std::string& msg) {
    if (c) c->err = msg;
    else g_create_err = msg;
    return code;
}
#define CU(ctx, call)                                                                                                  \
    do {                                                                                                               \
        cudaError_t e_ = (call);                                                                                       \
        if (e_ != cudaSuccess)                                                                                         \
            return fail(ctx, e_ == cudaErrorMemoryAllocation ? VK_ERR_OOM : VK_ERR_CUDA,                               \
                        std::string(#call) + ": " + cudaGetErrorString(e_));                                           \
    } while (0)

// ------------------------------------------------------------------------------------------------
// small kernels that do not depend on the math mode
// ------------------------------------------------------------------------------------------------
__global__ void k_reduce_chunks(const float* __restrict__ partial, uint32_t n_chunks, size_t n, float* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float acc = partial[i];
    for (uint32_t c = 1; c < n_chunks; ++c) acc += partial[(size_t)c * n + i]; // fixed order: deterministic
    out[i] = acc;
}
__global__ void k_finalize(const float* __restrict__ sum, float* __restrict__ rgb, size_t n, float spp) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) rgb[i] = sum[i] / spp; // c /= SAMPLES_PER_PIXEL as f32 (src/main.rs:196)
}
// Vec3::to_color (src/vec3.rs:54-61) of sum / spp, rows flipped to the PPM's top-down order (src/main.rs:209)
__global__ void k_to_color(const float* __restrict__ sum, uint8_t* __restrict__ out, uint32_t width, uint32_t height, float spp) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, n = (size_t)width * height * 3;
    if (i >= n) return;
    const size_t row = i / ((size_t)width * 3), col = i - row * (size_t)width * 3;
    const float c = sum[(size_t)(height - 1 - row) * width * 3 + col] / spp; // c /= SAMPLES_PER_PIXEL (main.rs:196)
    float v = sqrtf(c);
    v = v < 0.0f ? 0.0f : (v > 0.999f ? 0.999f : v); // Vec3::clamp keeps NaN
    out[i] = (uint8_t)__float2uint_rz(256.0f * v);   // `as u32`: NaN -> 0
}
// FP32 peak: 8 independent FFMA chains per thread, no memory traffic
__global__ void k_ffma_peak(float* out, int iters) {
    float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f, a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f,
          a7 = a0 + 7.f;
    const float m = 0.999f, b = 1e-3f;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            a0 = fmaf(a0, m, b); a1 = fmaf(a1, m, b); a2 = fmaf(a2, m, b); a3 = fmaf(a3, m, b);
            a4 = fmaf(a4, m, b); a5 = fmaf(a5, m, b); a6 = fmaf(a6, m, b); a7 = fmaf(a7, m, b);
        }
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
// L2 bandwidth: every block streams the same L2-resident buffer with 128-bit loads
__global__ void k_l2_read(const float4* __restrict__ buf, size_t n_vec, int reps, float* out) {
    float acc = 0.f;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int r = 0; r < reps; ++r)
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += stride) {
            const float4 v = __ldcg(&buf[i]);
            acc += v.x + v.y + v.z + v.w;
        }
    if (acc == 123.456f) out[0] = acc;
}

// ------------------------------------------------------------------------------------------------
// scene validation (host): everything the kernels assume is checked here, and anything the GPU
// path does not implement is refused with VK_ERR_UNSUPPORTED instead of being rendered wrongly.
// ------------------------------------------------------------------------------------------------
namespace {
inline float __uint_as_float_host(uint32_t u) {
    float f;
    std::memcpy(&f, &u, 4);
    return f;
}
struct Validator {
    const vk_scene_desc* d;
    std::string err;
    int code = VK_OK;
    std::vector<int> node_depth;       // 0 = unvisited
    std::vector<uint8_t> node_medium;  // subtree holds a ConstantMedium
    std::vector<uint8_t> node_inchain; // visited below an instance

    bool bad(int c, const std::string& m) {
        if (code == VK_OK) {
            code = c;
            err = m;
        }
        return false;
    }
    bool leaf_ok(vk_ref r) {
        const uint32_t i = VK_REF_INDEX(r);
        switch (VK_REF_TYPE(r)) {
        case VK_T_SPHERE: return i < d->n_spheres || bad(VK_ERR_INVALID, "sphere index out of range");
        case VK_T_MSPHERE: return i < d->n_mspheres || bad(VK_ERR_INVALID, "moving-sphere index out of range");
        case VK_T_RECT: return i < d->n_rects || bad(VK_ERR_INVALID, "rect index out of range");
        case VK_T_BOX: return i < d->n_boxes || bad(VK_ERR_INVALID, "box index out of range");
        default: return false;
        }
    }
    static bool is_leaf_type(uint32_t t) { return t == VK_T_SPHERE || t == VK_T_MSPHERE || t == VK_T_RECT || t == VK_T_BOX; }

    // returns the first non-wrapper ref of a chain starting at r (or NONE on error)
    vk_ref chain_end(vk_ref r) {
        int depth = 0;
        while (VK_REF_TYPE(r) == VK_T_XFORM) {
            if (VK_REF_INDEX(r) >= d->n_xforms) return bad(VK_ERR_INVALID, "xform index out of range"), VK_REF_NONE;
            if (++depth > VK_MAX_XFORM_DEPTH) return bad(VK_ERR_UNSUPPORTED, "wrapper chain deeper than VK_MAX_XFORM_DEPTH"), VK_REF_NONE;
            if (d->xforms[VK_REF_INDEX(r)].kind > VK_X_FLIP) return bad(VK_ERR_INVALID, "bad xform kind"), VK_REF_NONE;
            r = d->xforms[VK_REF_INDEX(r)].child;
        }
        return r;
    }
    bool medium_ok(vk_ref r, bool in_chain) {
        if (VK_REF_INDEX(r) >= d->n_media) return bad(VK_ERR_INVALID, "medium index out of range");
        const vk_medium& m = d->media[VK_REF_INDEX(r)];
        if (m.mat >= d->n_materials) return bad(VK_ERR_INVALID, "medium material out of range");
        vk_ref b = m.boundary;
        if (VK_REF_TYPE(b) == VK_T_XFORM) {
            if (in_chain) return bad(VK_ERR_UNSUPPORTED, "ConstantMedium with a transformed boundary inside another instance");
            b = chain_end(b);
            if (code != VK_OK) return false;
        }
        if (!is_leaf_type(VK_REF_TYPE(b)))
            return bad(VK_ERR_UNSUPPORTED, "ConstantMedium boundary must be a sphere, rect or box (optionally translated/rotated)");
        return leaf_ok(b);
    }
    // depth of the traversal stack needed below r; also fills node_medium
    int visit(vk_ref r, bool in_chain, bool& has_medium) {
        has_medium = false;
        const uint32_t t = VK_REF_TYPE(r), i = VK_REF_INDEX(r);
        if (t == VK_T_NODE) {
            if (i >= d->n_nodes) return bad(VK_ERR_INVALID, "node index out of range"), 0;
            if (node_depth[i] == -1) return bad(VK_ERR_INVALID, "cycle in the BVH"), 0;
            if (node_depth[i] > 0) {
                if (in_chain && !node_inchain[i]) node_inchain[i] = 1; // (shape already validated; chain rule rechecked below)
                has_medium = node_medium[i];
                return node_depth[i];
            }
            node_depth[i] = -1;
            bool ml = false, mr = false;
            const int dl = visit(d->nodes[i].left, in_chain, ml);
            const int dr = d->nodes[i].right == d->nodes[i].left ? dl : visit(d->nodes[i].right, in_chain, mr);
            if (d->nodes[i].right == d->nodes[i].left) mr = ml;
            if (code != VK_OK) return 0;
            if (d->nodes[i].right == d->nodes[i].left && ml && VK_REF_TYPE(chain_end(d->nodes[i].left)) != VK_T_MEDIUM)
                return bad(VK_ERR_UNSUPPORTED, "single-object BVH leaf holding a medium inside a nested BVH"), 0;
            node_medium[i] = ml || mr;
            node_inchain[i] = in_chain;
            has_medium = node_medium[i];
            node_depth[i] = 1 + (dl > dr ? dl : dr);
            return node_depth[i];
        }
        if (t == VK_T_XFORM) {
            if (in_chain) return bad(VK_ERR_UNSUPPORTED, "nested instances (a transform below another transform's BVH)"), 0;
            const vk_ref end = chain_end(r);
            if (code != VK_OK) return 0;
            if (VK_REF_TYPE(end) == VK_T_NODE) return 2 + visit(end, true, has_medium);
            if (VK_REF_TYPE(end) == VK_T_MEDIUM) {
                has_medium = true;
                medium_ok(end, true);
                return 1;
            }
            if (!is_leaf_type(VK_REF_TYPE(end))) return bad(VK_ERR_INVALID, "bad reference below a transform"), 0;
            leaf_ok(end);
            return 1;
        }
        if (t == VK_T_MEDIUM) {
            has_medium = true;
            medium_ok(r, in_chain);
            return 1;
        }
        if (is_leaf_type(t)) {
            leaf_ok(r);
            return 1;
        }
        return bad(VK_ERR_INVALID, "bad hittable reference"), 0;
    }
    bool tex_has_image(uint32_t ti, int depth) {
        if (ti >= d->n_textures || depth > 16) return false;
        const vk_texture& t = d->textures[ti];
        if (t.type == VK_TEX_IMAGE) return true;
        if (t.type == VK_TEX_CHECKER) return tex_has_image(t.checker.odd, depth + 1) || tex_has_image(t.checker.even, depth + 1);
        return false;
    }
    bool run() {
        if (!d || d->api_version != VK_API_VERSION) return bad(VK_ERR_INVALID, "scene description: wrong api_version");
        for (uint32_t i = 0; i < d->n_textures; ++i) {
            const vk_texture& t = d->textures[i];
            if (t.type > VK_TEX_NOISE) return bad(VK_ERR_INVALID, "bad texture type");
            if (t.type == VK_TEX_CHECKER && (t.checker.odd >= d->n_textures || t.checker.even >= d->n_textures))
                return bad(VK_ERR_INVALID, "checker child out of range");
            if (t.type == VK_TEX_IMAGE && (t.image.width == 0 || t.image.height == 0 ||
                                           (uint64_t)t.image.texel_offset + (uint64_t)t.image.width * t.image.height * 3 > d->n_texel_bytes))
                return bad(VK_ERR_INVALID, "image texture outside the texel pool");
            if (t.type == VK_TEX_NOISE && t.noise.perlin >= d->n_perlins) return bad(VK_ERR_INVALID, "perlin index out of range");
        }
        for (uint32_t i = 0; i < d->n_materials; ++i) {
            const vk_material& m = d->materials[i];
            if (m.type > VK_M_SPECDIFFUSE) return bad(VK_ERR_INVALID, "bad material type");
            if (m.type == VK_M_SPECDIFFUSE) {
                if (m.tex >= d->n_materials || m.aux >= d->n_materials) return bad(VK_ERR_INVALID, "SpecDiffuse child out of range");
                if (d->materials[m.tex].type == VK_M_SPECDIFFUSE || d->materials[m.aux].type == VK_M_SPECDIFFUSE)
                    return bad(VK_ERR_UNSUPPORTED, "nested SpecDiffuse");
            } else if (m.type != VK_M_DIELECTRIC && m.tex >= d->n_textures)
                return bad(VK_ERR_INVALID, "material texture out of range");
        }
        for (uint32_t i = 0; i < d->n_spheres; ++i)
            if (d->sphere_mat[i] >= d->n_materials) return bad(VK_ERR_INVALID, "sphere material out of range");
        for (uint32_t i = 0; i < d->n_mspheres; ++i)
            if (d->mspheres[i].mat >= d->n_materials) return bad(VK_ERR_INVALID, "moving-sphere material out of range");
        for (uint32_t i = 0; i < d->n_rects; ++i) {
            const uint32_t ax = d->rects[i].axes;
            if (d->rects[i].mat >= d->n_materials) return bad(VK_ERR_INVALID, "rect material out of range");
            if ((ax & 3) > 2 || ((ax >> 2) & 3) > 2 || ((ax >> 4) & 3) > 2) return bad(VK_ERR_INVALID, "rect axis out of range");
        }
        for (uint32_t i = 0; i < d->n_boxes; ++i)
            if (d->boxes[i].mat >= d->n_materials) return bad(VK_ERR_INVALID, "box material out of range");
        if (d->n_nodes > VKD_INDEX(0xFFFFFFFFu) || d->n_spheres > VKD_INDEX(0xFFFFFFFFu)) return bad(VK_ERR_UNSUPPORTED, "too many primitives");
        node_depth.assign(d->n_nodes, 0);
        node_medium.assign(d->n_nodes, 0);
        node_inchain.assign(d->n_nodes, 0);
        bool hm = false;
        const int depth = visit(d->root, false, hm);
        if (code != VK_OK) return false;
        if (depth + 2 > VKD_STACK) return bad(VK_ERR_UNSUPPORTED, "BVH deeper than the traversal stack");
        // an empty light list is accepted here: HEAD's integrator is refused at render time (the reference
        // panics, src/hittable.rs:431), the legacy integrator (VK_FLAG_LEGACY_SCATTER) does not use it
        for (uint32_t i = 0; i < d->n_lights; ++i) {
            const vk_ref l = d->lights[i];
            const uint32_t t = VK_REF_TYPE(l), ix = VK_REF_INDEX(l);
            const uint32_t lim = t == VK_T_NODE ? d->n_nodes : t == VK_T_SPHERE ? d->n_spheres : t == VK_T_MSPHERE ? d->n_mspheres
                               : t == VK_T_RECT ? d->n_rects : t == VK_T_BOX ? d->n_boxes : t == VK_T_XFORM ? d->n_xforms
                               : t == VK_T_MEDIUM ? d->n_media : 0;
            if (ix >= lim) return bad(VK_ERR_INVALID, "light reference out of range");
        }
        return true;
    }
};

// Unroll the reference's traversal into the typed batches of FlatProgram; false if the scene does not
// fit (then the BVH is traversed).
struct FlatBuilder {
    const vk_scene_desc* d;
    struct Seg {
        std::vector<FlatOp> ops;
        std::vector<std::pair<FlatRect, FlatHit>> rects[6];
        std::vector<std::pair<FlatSphere, FlatHit>> sph, msph;
        std::vector<FlatHit> med;
        std::vector<uint32_t> bvh; // roots of homogeneous subtrees kept as BVHs (hybrid program)
        uint32_t inst = 0;
    };
    std::vector<Seg> segs;
    // hybrid mode: a subtree whose leaves are all of ONE plain primitive kind (no wrapper, no medium) and
    // that holds more than VKF_SUBTREE_MIN of them stays a BVH; only the mixed top of the tree is unrolled
    bool hybrid = false;
    struct NodeInfo {
        uint32_t leaves = 0;
        uint8_t kind = 0; // VK_T_* shared by every leaf below, 0xFF = mixed / wrapper / medium
        bool done = false;
    };
    std::vector<NodeInfo> info;
    NodeInfo classify(vk_ref r) {
        NodeInfo out;
        const uint32_t t = VK_REF_TYPE(r);
        if (t != VK_T_NODE) {
            out.leaves = 1;
            out.kind = (t == VK_T_SPHERE || t == VK_T_RECT || t == VK_T_BOX) ? (uint8_t)t : (uint8_t)0xFF;
            return out;
        }
        NodeInfo& me = info[VK_REF_INDEX(r)];
        if (me.done) return me;
        const vk_node& n = d->nodes[VK_REF_INDEX(r)];
        const NodeInfo a = classify(n.left), b = n.right == n.left ? a : classify(n.right);
        me.leaves = a.leaves + (n.right == n.left ? 0 : b.leaves);
        me.kind = (a.kind == b.kind) ? a.kind : (uint8_t)0xFF;
        me.done = true;
        return me;
    }

    bool rect(Seg& g, float c0, float c1, float d0, float d1, float k, uint32_t axes, vk_ref ref, uint32_t face, bool box_side) {
        const uint32_t a0 = axes & 3u, a1 = (axes >> 2) & 3u, a2 = (axes >> 4) & 3u;
        // the canonical axis order of Rect::XYRect/XZRect/YZRect (src/hittable.rs:214-226) is assumed
        const bool canonical = (a2 == 2 && a0 == 0 && a1 == 1) || (a2 == 1 && a0 == 0 && a1 == 2) || (a2 == 0 && a0 == 1 && a1 == 2);
        if (!canonical) return false;
        FlatRect e{};
        e.bounds = make_float4(c0, c1, d0, d1);
        e.k = k;
        FlatHit h{ref, g.inst, face, 0u};
        g.rects[(a2 == 2 ? 0 : (a2 == 1 ? 1 : 2)) + (box_side ? 3 : 0)].push_back({e, h});
        return true;
    }
    size_t n_emitted = 0; // entries so far: a scene that cannot fit is abandoned early, not walked to the end
    bool emit(vk_ref ref, size_t si, uint32_t dup) {
        if (++n_emitted > 4 * (VKF_MAX_RECTS + VKF_MAX_SPHERES + VKF_MAX_MEDIA + VKF_MAX_BVH + VKF_MAX_OPS)) return false;
        const uint32_t i = VK_REF_INDEX(ref);
        switch (VK_REF_TYPE(ref)) {
        case VK_T_NODE: {
            if (hybrid) {
                const NodeInfo ni = classify(ref);
                if (ni.kind != 0xFF && ni.leaves > 8) { // homogeneous and worth a BVH: keep it as one entry
                    segs[si].bvh.push_back(ref);
                    return true;
                }
            }
            const vk_node& n = d->nodes[i];
            if (!emit(n.left, si, dup)) return false;
            if (n.right != n.left) return emit(n.right, si, dup);
            vk_ref end = n.left; // single-object leaf: the second visit only matters for a medium
            while (VK_REF_TYPE(end) == VK_T_XFORM) end = d->xforms[VK_REF_INDEX(end)].child;
            return VK_REF_TYPE(end) == VK_T_MEDIUM ? emit(n.left, si, VKD_DUP) : true;
        }
        case VK_T_SPHERE: {
            FlatSphere e{};
            e.a = make_float4(d->spheres[i].center[0], d->spheres[i].center[1], d->spheres[i].center[2], d->spheres[i].radius);
            segs[si].sph.push_back({e, FlatHit{ref, segs[si].inst, 0u, 0u}});
            return true;
        }
        case VK_T_MSPHERE: {
            const vk_msphere& m = d->mspheres[i];
            FlatSphere e{};
            e.a = make_float4(m.center0[0], m.center0[1], m.center0[2], m.radius);
            e.b = make_float4(m.center1[0], m.center1[1], m.center1[2], m.time0);
            e.time1 = m.time1;
            segs[si].msph.push_back({e, FlatHit{ref, segs[si].inst, 0u, 0u}});
            return true;
        }
        case VK_T_RECT: {
            const vk_rect& r = d->rects[i];
            return rect(segs[si], r.c0, r.c1, r.d0, r.d1, r.k, r.axes, ref, 0, false);
        }
        case VK_T_BOX: { // the six sides in Boxy::new order (src/hittable.rs:325-353)
            const vk_box& b = d->boxes[i];
            const float* mn = b.box_min;
            const float* mx = b.box_max;
            const uint32_t XY = 0u | (1u << 2) | (2u << 4), XZ = 0u | (2u << 2) | (1u << 4), YZ = 1u | (2u << 2) | (0u << 4);
            Seg& g = segs[si];
            return rect(g, mn[0], mx[0], mn[1], mx[1], mx[2], XY, ref, 0, true) && rect(g, mn[0], mx[0], mn[1], mx[1], mn[2], XY, ref, 1, true) &&
                   rect(g, mn[0], mx[0], mn[2], mx[2], mx[1], XZ, ref, 2, true) && rect(g, mn[0], mx[0], mn[2], mx[2], mn[1], XZ, ref, 3, true) &&
                   rect(g, mn[1], mx[1], mn[2], mx[2], mx[0], YZ, ref, 4, true) && rect(g, mn[1], mx[1], mn[2], mx[2], mn[0], YZ, ref, 5, true);
        }
        case VK_T_MEDIUM:
            segs[si].med.push_back(FlatHit{ref | dup, segs[si].inst, 0u, 0u});
            return true;
        case VK_T_XFORM: {
            if (si != 0) return false; // nested instances are refused by the validator anyway
            Seg g;
            g.inst = ref; // instance id = outermost wrapper
            vk_ref r = ref;
            while (VK_REF_TYPE(r) == VK_T_XFORM) {
                const vk_xform& x = d->xforms[VK_REF_INDEX(r)];
                if (x.kind == VK_X_TRANSLATE) g.ops.push_back(FlatOp{VKF_OP_TRANSLATE, x.a, x.b, x.c});
                else if (x.kind != VK_X_FLIP) // FlipFace leaves the ray alone; the flip happens in resolve_hit
                    g.ops.push_back(FlatOp{x.kind == VK_X_ROTATE_X ? (uint32_t)VKF_OP_ROTX : (x.kind == VK_X_ROTATE_Y ? (uint32_t)VKF_OP_ROTY : (uint32_t)VKF_OP_ROTZ), x.a, x.b, 0.f});
                r = x.child;
            }
            segs.push_back(std::move(g));
            return emit(r, segs.size() - 1, dup);
        }
        default: return false;
        }
    }
    bool build(FlatProgram* P, bool hybrid_mode) {
        *P = FlatProgram{};
        hybrid = hybrid_mode;
        n_emitted = 0;
        info.assign(hybrid ? d->n_nodes : 0, NodeInfo{});
        segs.clear();
        segs.emplace_back();
        if (!emit(d->root, 0, 0)) return false;
        if (segs.size() > VKF_MAX_SEGS) return false;
        uint32_t n_ops = 0, n_rects = 0, n_sph = 0, n_hits = 0, n_med = 0, n_bvh = 0;
        for (size_t s = 0; s < segs.size(); ++s) {
            Seg& g = segs[s];
            FlatSeg& o = P->segs[s];
            if (n_ops + g.ops.size() > VKF_MAX_OPS) return false;
            o.op0 = (uint8_t)n_ops;
            for (const FlatOp& op : g.ops) P->ops[n_ops++] = op;
            o.op1 = (uint8_t)n_ops;
            for (int k = 0; k < 6; ++k) {
                if (n_rects + g.rects[k].size() > VKF_MAX_RECTS) return false;
                o.rect0[k] = (uint8_t)n_rects;
                for (auto& e : g.rects[k]) {
                    e.first.hit = n_hits;
                    P->hits[n_hits++] = e.second;
                    P->rects[n_rects++] = e.first;
                }
                o.rect1[k] = (uint8_t)n_rects;
            }
            if (n_sph + g.sph.size() + g.msph.size() > VKF_MAX_SPHERES) return false;
            o.sph0 = (uint8_t)n_sph;
            for (auto& e : g.sph) {
                e.first.hit = n_hits;
                P->hits[n_hits++] = e.second;
                P->spheres[n_sph++] = e.first;
            }
            o.sph1 = o.msph0 = (uint8_t)n_sph;
            for (auto& e : g.msph) {
                e.first.hit = n_hits;
                P->hits[n_hits++] = e.second;
                P->spheres[n_sph++] = e.first;
            }
            o.msph1 = (uint8_t)n_sph;
            if (n_med + g.med.size() > VKF_MAX_MEDIA) return false;
            o.med0 = (uint8_t)n_hits;
            for (const FlatHit& h : g.med) {
                P->hits[n_hits++] = h;
                ++n_med;
            }
            o.med1 = (uint8_t)n_hits;
            if (n_bvh + g.bvh.size() > VKF_MAX_BVH) return false;
            o.bvh0 = (uint8_t)n_bvh;
            for (uint32_t r : g.bvh) P->bvh[n_bvh++] = r;
            o.bvh1 = (uint8_t)n_bvh;
            P->seg_inst[s] = g.inst;
        }
        P->n_bvh = n_bvh;
        for (uint32_t h = 0; h < n_hits; ++h) { // shading class of each entry's material
            const vk_ref pr = P->hits[h].prim & ~VKD_DUP;
            const uint32_t i = VK_REF_INDEX(pr);
            uint32_t mat = 0;
            switch (VK_REF_TYPE(pr)) {
            case VK_T_SPHERE: mat = d->sphere_mat[i]; break;
            case VK_T_MSPHERE: mat = d->mspheres[i].mat; break;
            case VK_T_RECT: mat = d->rects[i].mat; break;
            case VK_T_BOX: mat = d->boxes[i].mat; break;
            case VK_T_MEDIUM: mat = d->media[i].mat; break;
            default: return false;
            }
            const uint32_t t = d->materials[mat].type;
            P->hits[h].cls = t == VK_M_DIFFUSE_LIGHT ? 0u : t == VK_M_DIELECTRIC ? 1u : t == VK_M_METAL ? 2u : 3u;
        }
        P->n_segs = (uint32_t)segs.size();
        P->n = n_hits + n_bvh;
        return P->n > 0;
    }
};

// Host-side re-layout of a validated scene (no device needed): the node array with single-object
// leaves resolved, the 4-wide nodes, the flat program and the "simple scene" test.  vk_scene_upload
// copies the results to the device; vk_scene_check reports them so that CPU tests can cover this code.
struct Relayout {
    std::vector<vk_node> nodes;
    std::vector<float4> wnodes;
    FlatProgram flat{};
    uint32_t levels_world = 0, levels_sub = 0, n_wide = 0, stack_need = 0;
    bool simple = false, has_specdiffuse = false;
    // returns nullptr or the reason the scene is unsupported
    const char* run(const vk_scene_desc* d) {
    // GPU-side re-layout of the node array: a single-object leaf (left == right) is tested twice
        // by the reference; that only matters for a ConstantMedium (two free-flight draws), so the
        // second visit is kept (flagged) only there and dropped for deterministic primitives.
        nodes.assign(d->nodes, d->nodes + d->n_nodes);
        for (uint32_t i = 0; i < d->n_nodes; ++i)
            if (nodes[i].left == nodes[i].right) {
                vk_ref end = nodes[i].left;
                while (VK_REF_TYPE(end) == VK_T_XFORM) end = d->xforms[VK_REF_INDEX(end)].child;
                nodes[i].right = VK_REF_TYPE(end) == VK_T_MEDIUM ? (nodes[i].left | VKD_DUP) : VK_REF_NONE;
            }
        // 4-wide nodes from the reference's binary tree (see DScene): start from a node's two children and
        // keep opening the inner child with the largest surface area until four slots are filled.
        wnodes.assign((size_t)d->n_nodes * 8, make_float4(0, 0, 0, 0));
        {
            struct Slot {
                vk_ref ref;
                float mn[3], mx[3];
            };
            std::vector<uint8_t> built(d->n_nodes, 0);
            std::vector<uint32_t> todo, level(d->n_nodes, 0); // level: 4-wide levels above the node inside its BVH
            levels_world = levels_sub = 0;
            bool in_sub = false;
            auto want = [&](vk_ref r, uint32_t lvl) {
                if (VK_REF_TYPE(r) == VK_T_NODE && !built[VK_REF_INDEX(r)]) {
                    built[VK_REF_INDEX(r)] = 1;
                    level[VK_REF_INDEX(r)] = lvl;
                    todo.push_back(VK_REF_INDEX(r));
                    uint32_t& top = in_sub ? levels_sub : levels_world;
                    if (lvl + 1 > top) top = lvl + 1;
                }
            };
            auto child_slot = [&](vk_ref r, const vk_node& parent) { // a node child brings its own box, a primitive its parent's
                const vk_node& b = VK_REF_TYPE(r) == VK_T_NODE ? d->nodes[VK_REF_INDEX(r)] : parent;
                Slot s{r, {b.bb_min[0], b.bb_min[1], b.bb_min[2]}, {b.bb_max[0], b.bb_max[1], b.bb_max[2]}};
                return s;
            };
            auto area = [](const Slot& s) {
                const float x = s.mx[0] - s.mn[0], y = s.mx[1] - s.mn[1], z = s.mx[2] - s.mn[2];
                return x * y + y * z + z * x;
            };
            // the world's BVH first, then the instanced sub-BVHs (an instance is never nested: see Validator)
            for (int pass = 0; pass < 2; ++pass) {
            in_sub = pass == 1;
            if (pass == 0) want(d->root, 0);
            else
                for (uint32_t i = 0; i < d->n_xforms; ++i) want(d->xforms[i].child, 0);
            while (!todo.empty()) {
                const uint32_t ni = todo.back();
                todo.pop_back();
                Slot slots[6]; // never more than four after a round; two are added before one is removed
                size_t n_slots = 0;
                auto add_children = [&](uint32_t n) {
                    if (nodes[n].left != VK_REF_NONE) slots[n_slots++] = child_slot(nodes[n].left, d->nodes[n]);
                    if (nodes[n].right != VK_REF_NONE) slots[n_slots++] = child_slot(nodes[n].right, d->nodes[n]);
                };
                add_children(ni);
                while (n_slots < 4) {
                    int best = -1;
                    for (size_t k = 0; k < n_slots; ++k)
                        if (VK_REF_TYPE(slots[k].ref) == VK_T_NODE) {
                            const uint32_t n = VK_REF_INDEX(slots[k].ref);
                            const size_t kids = (nodes[n].left != VK_REF_NONE) + (nodes[n].right != VK_REF_NONE);
                            if (n_slots - 1 + kids > 4) continue;
                            if (best < 0 || area(slots[k]) > area(slots[best])) best = (int)k;
                        }
                    if (best < 0) break;
                    const uint32_t n = VK_REF_INDEX(slots[best].ref);
                    for (size_t k = (size_t)best; k + 1 < n_slots; ++k) slots[k] = slots[k + 1];
                    --n_slots;
                    add_children(n);
                }
                ++n_wide;
                float4* q = &wnodes[(size_t)ni * 8];
                float* f = reinterpret_cast<float*>(q);
                for (size_t k = 0; k < 4; ++k) {
                    const bool have = k < n_slots;
                    for (int ax = 0; ax < 3; ++ax) {
                        f[(2 * ax) * 4 + k] = have ? slots[k].mn[ax] : 0.f;
                        f[(2 * ax + 1) * 4 + k] = have ? slots[k].mx[ax] : 0.f;
                    }
                    f[6 * 4 + k] = __uint_as_float_host(have ? slots[k].ref : VK_REF_NONE);
                    if (have) want(slots[k].ref, level[ni] + 1);
                }
            }
            }
            // a visit pushes at most three siblings; an instance adds its exit marker
            if (3 * levels_world + 1 + 3 * levels_sub + 2 > VKD_STACK)
                return "BVH deeper than the traversal stack";
        }
        stack_need = 3 * levels_world + 1 + 3 * levels_sub + 2;
        // the whole scene as typed batches if it is small; else, for a heterogeneous scene (wrappers or media
        // present), the mixed top of the tree as batches and its homogeneous subtrees as BVH entries
        FlatBuilder fb;
        fb.d = d;
        if (!fb.build(&flat, false)) {
            const bool heterogeneous = d->n_media > 0 || d->n_xforms > 0;
            // Off by default: measured on the final scene (800x800x64) the hybrid program renders in 52.5 ms
            // against 50.0 ms for the plain 4-wide BVH -- the flat top tests all nine loose objects (two
            // media included) for every ray where the BVH culls some, and the subtree traversals keep their
            // divergence.  VECCHIO_HYBRID=1 enables it (parity-tested: same image bit for bit).
            if (!(heterogeneous && std::getenv("VECCHIO_HYBRID") && fb.build(&flat, true) && flat.n_bvh > 0)) flat = FlatProgram{};
        }
        has_specdiffuse = false;
        for (uint32_t i = 0; i < d->n_materials; ++i) has_specdiffuse |= d->materials[i].type == VK_M_SPECDIFFUSE;
        // simple: solid textures only; Lambertian / Dielectric / DiffuseLight / Isotropic only; no moving sphere;
        // exactly one light, an unflipped Rect
        simple = d->n_mspheres == 0 && d->n_lights == 1 && VK_REF_TYPE(d->lights[0]) == VK_T_RECT &&
                 !(d->rects[VK_REF_INDEX(d->lights[0])].axes & VK_RECT_FLIP);
        for (uint32_t i = 0; i < d->n_textures && simple; ++i) simple = d->textures[i].type == VK_TEX_SOLID;
        for (uint32_t i = 0; i < d->n_materials && simple; ++i)
            simple = d->materials[i].type == VK_M_LAMBERTIAN || d->materials[i].type == VK_M_DIELECTRIC ||
                     d->materials[i].type == VK_M_DIFFUSE_LIGHT || d->materials[i].type == VK_M_ISOTROPIC;
        return nullptr;
    }
};

template <class T> int upload(vk_ctx* c, const T* src, size_t n, const T** dst) {
    *dst = nullptr;
    if (n == 0) n = 1; // keep pointers valid
    void* p = nullptr;
    CU(c, cudaMalloc(&p, n * sizeof(T)));
    c->scene_allocs.push_back(p);
    if (src) CU(c, cudaMemcpyAsync(p, src, n * sizeof(T), cudaMemcpyHostToDevice, c->stream));
    *dst = (const T*)p;
    return VK_OK;
}
void free_scene(vk_ctx* c) {
    for (void* p : c->scene_allocs) cudaFree(p);
    c->scene_allocs.clear();
    c->has_scene = false;
}
int ensure(vk_ctx* c, float** buf, size_t* have, size_t want) {
    if (*have >= want) return VK_OK;
    if (*buf) cudaFree(*buf);
    *buf = nullptr;
    *have = 0;
    CU(c, cudaMalloc((void**)buf, want * sizeof(float)));
    *have = want;
    return VK_OK;
}
DCamera to_dcam(const vk_camera* k) {
    DCamera c;
    auto v3 = [](const float* p) { return make_float3(p[0], p[1], p[2]); };
    c.origin = v3(k->origin);
    c.lower_left_corner = v3(k->lower_left_corner);
    c.horizontal = v3(k->horizontal);
    c.vertical = v3(k->vertical);
    c.u = v3(k->u);
    c.v = v3(k->v);
    c.lens_radius = k->lens_radius;
    c.time0 = k->time0;
    c.time1 = k->time1;
    return c;
}
} // namespace

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

int vk_create(int device, vk_ctx** out) {
    if (!out) return fail(nullptr, VK_ERR_INVALID, "vk_create: null out pointer");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(nullptr, VK_ERR_NO_DEVICE, std::string("vk_create: no CUDA device (") + cudaGetErrorString(e) +
                                                   "); this library has no CPU fallback");
    if (device < 0 || device >= n) return fail(nullptr, VK_ERR_INVALID, "vk_create: device index out of range");
    vk_ctx* c = new vk_ctx;
    c->device = device;
#define CUC(call)                                                                                                      \
    do {                                                                                                               \
        cudaError_t e_ = (call);                                                                                       \
        if (e_ != cudaSuccess) {                                                                                       \
            g_create_err = std::string(#call) + ": " + cudaGetErrorString(e_);                                         \
            delete c;                                                                                                  \
            return VK_ERR_CUDA;                                                                                        \
        }                                                                                                              \
    } while (0)
    CUC(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUC(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        g_create_err = std::string("vk_create: device '") + prop.name + "' is sm_" + std::to_string(prop.major * 10 + prop.minor) +
                       "; this library is built for sm_100a only";
        delete c;
        return VK_ERR_NO_DEVICE;
    }
    c->sm_count = prop.multiProcessorCount;
    c->clock_khz = prop.clockRate;
    std::snprintf(c->name, sizeof(c->name), "%.100s", prop.name);
    CUC(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    c->stream = c->own_stream;
    CUC(cudaEventCreate(&c->ev0));
    CUC(cudaEventCreate(&c->ev1));
    CUC(cudaEventCreate(&c->ev2));
    CUC(cudaMalloc((void**)&c->counters, 8 * sizeof(unsigned long long)));
    CUC(cudaMemset(c->counters, 0, 8 * sizeof(unsigned long long)));
    CUC(cudaMalloc((void**)&c->debug, VK_DEBUG_CTAS * 4 * sizeof(unsigned long long)));
    CUC(cudaMemset(c->debug, 0, VK_DEBUG_CTAS * 4 * sizeof(unsigned long long)));
#undef CUC
    *out = c;
    return VK_OK;
}

void vk_destroy(vk_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    free_scene(c);
    if (c->partial) cudaFree(c->partial);
    if (c->wf_block) cudaFree(c->wf_block);
    if (c->wf_host_counts) cudaFreeHost(c->wf_host_counts);
    if (c->frame) cudaFree(c->frame);
    if (c->pinned) cudaFreeHost(c->pinned);
    if (c->counters) cudaFree(c->counters);
    if (c->debug) cudaFree(c->debug);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->ev2) cudaEventDestroy(c->ev2);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
}

const char* vk_last_error(const vk_ctx* c) { return c ? c->err.c_str() : g_create_err.c_str(); }

int vk_device_info(vk_ctx* c, int* sm_count, int* clock_khz, char* name, size_t name_len) {
    if (!c) return VK_ERR_INVALID;
    if (sm_count) *sm_count = c->sm_count;
    if (clock_khz) *clock_khz = c->clock_khz;
    if (name && name_len) std::snprintf(name, name_len, "%s", c->name);
    return VK_OK;
}

int vk_scene_check(const vk_scene_desc* d, vk_scene_info* info, char* err, size_t err_len) {
    auto report = [&](int code, const std::string& msg) {
        if (err && err_len) std::snprintf(err, err_len, "%s", msg.c_str());
        return code;
    };
    if (!d) return report(VK_ERR_INVALID, "null scene");
    Validator v;
    v.d = d;
    if (!v.run()) return report(v.code, v.err);
    Relayout R;
    if (const char* why = R.run(d)) return report(VK_ERR_UNSUPPORTED, why);
    if (info) {
        info->flat_entries = R.flat.n;
        info->flat_segments = R.flat.n ? R.flat.n_segs : 0;
        info->flat_subtrees = R.flat.n ? R.flat.n_bvh : 0;
        info->simple = R.simple ? 1u : 0u;
        info->wide_nodes = R.n_wide;
        info->stack_need = R.stack_need;
        info->wide_levels_world = R.levels_world;
        info->wide_levels_instance = R.levels_sub;
        info->dynamic_megakernel = d->n_nodes >= 65536u ? 1u : 0u;
    }
    if (err && err_len) err[0] = 0;
    return VK_OK;
}

int vk_scene_upload(vk_ctx* c, const vk_scene_desc* d) {
    if (!c) return VK_ERR_INVALID;
    if (!d) return fail(c, VK_ERR_INVALID, "vk_scene_upload: null scene");
    Validator v;
    v.d = d;
    if (!v.run()) return fail(c, v.code, "vk_scene_upload: " + v.err);
    CU(c, cudaSetDevice(c->device));
    free_scene(c);

    Relayout R;
    if (const char* why = R.run(d)) return fail(c, VK_ERR_UNSUPPORTED, std::string("vk_scene_upload: ") + why);
    std::vector<vk_node>& nodes = R.nodes;
    std::vector<float4>& wnodes = R.wnodes;
    std::vector<vk_material> mats(d->materials, d->materials + d->n_materials);
    for (uint32_t i = 0; i < d->n_materials; ++i) {
        bool uv;
        if (mats[i].type == VK_M_SPECDIFFUSE)
            uv = (d->materials[mats[i].tex].type != VK_M_DIELECTRIC && v.tex_has_image(d->materials[mats[i].tex].tex, 0)) ||
                 (d->materials[mats[i].aux].type != VK_M_DIELECTRIC && v.tex_has_image(d->materials[mats[i].aux].tex, 0));
        else
            uv = mats[i].type != VK_M_DIELECTRIC && v.tex_has_image(mats[i].tex, 0);
        if (uv) mats[i].aux |= VKD_MAT_NEEDS_UV;
    }
    std::vector<float4> pvec((size_t)d->n_perlins * 256);
    std::vector<uint8_t> pperm((size_t)d->n_perlins * 768);
    for (uint32_t p = 0; p < d->n_perlins; ++p) {
        for (int k = 0; k < 256; ++k)
            pvec[(size_t)p * 256 + k] = make_float4(d->perlins[p].ranvec[k][0], d->perlins[p].ranvec[k][1], d->perlins[p].ranvec[k][2], 0.f);
        std::memcpy(&pperm[(size_t)p * 768], d->perlins[p].perm_x, 256);
        std::memcpy(&pperm[(size_t)p * 768 + 256], d->perlins[p].perm_y, 256);
        std::memcpy(&pperm[(size_t)p * 768 + 512], d->perlins[p].perm_z, 256);
    }

    DScene s{};
    int rc;
#define UP(field, type, src, n)                                                                                        \
    if ((rc = upload<type>(c, (const type*)(src), (n), (const type**)&s.field)) != VK_OK) {                            \
        free_scene(c);                                                                                                 \
        return rc;                                                                                                     \
    }
    UP(wnodes, float4, wnodes.data(), wnodes.size())
    UP(nodes, float4, nodes.data(), (size_t)d->n_nodes * 2)
    UP(spheres, float4, d->spheres, d->n_spheres)
    UP(sphere_mat, uint32_t, d->sphere_mat, d->n_spheres)
    UP(mspheres, float4, d->mspheres, (size_t)d->n_mspheres * 3)
    UP(rects, float4, d->rects, (size_t)d->n_rects * 2)
    UP(boxes, float4, d->boxes, (size_t)d->n_boxes * 2)
    UP(xforms, float4, d->xforms, (size_t)d->n_xforms * 2)
    UP(media, float4, d->media, d->n_media)
    UP(lights, uint32_t, d->lights, d->n_lights)
    UP(materials, uint4, mats.data(), d->n_materials)
    UP(textures, uint4, d->textures, d->n_textures)
    UP(texels, uint8_t, d->texels, (size_t)d->n_texel_bytes)
    UP(perlin_vec, float4, pvec.data(), pvec.size())
    UP(perlin_perm, uint8_t, pperm.data(), pperm.size())
#undef UP
    s.root = d->root;
    s.n_lights = d->n_lights;
    s.has_media = d->n_media > 0;
    CU(c, cudaStreamSynchronize(c->stream)); // host staging vectors die at return
    c->scene = s;
    c->n_nodes = d->n_nodes;
    c->has_specdiffuse = R.has_specdiffuse;
    c->simple_scene = R.simple;
    c->flat = R.flat;
    c->has_scene = true;
    return VK_OK;
}

extern "C" int vk_flush_stats(vk_ctx* c, vk_stats* stats);

// Slot pool of the wavefront variant.  One allocation, carved into the arrays of WfState.  Default
// 2^19 slots: 96 B (112 B with sum of squares) per slot = 50 MB, resident in the 126 MB L2, and
// 3-4 full waves of 256-thread CTAs per launch.  VECCHIO_WF_SLOTS overrides (tuning sweeps).
static int wf_ensure(vk_ctx* c, bool want_sumsq) {
    uint32_t n = 1u << 19;
    if (const char* e = std::getenv("VECCHIO_WF_SLOTS")) {
        const long v = std::atol(e);
        if (v >= 1024 && v <= (1l << 26)) n = (uint32_t)v;
    }
    n = (n + 255u) & ~255u;
    if (c->wf_block && c->wf.n_slots == n && (c->wf_has_sumsq || !want_sumsq)) return VK_OK;
    if (c->wf_block) cudaFree(c->wf_block);
    c->wf_block = nullptr;
    c->wf = WfState{};
    const size_t per_slot = 16 * (6 + (want_sumsq ? 1 : 0)) + 4 * VKW_CLASSES;
    const size_t tail = 256; // qcount (2 sets) + unit_head
    CU(c, cudaMalloc(&c->wf_block, per_slot * n + tail));
    char* p = (char*)c->wf_block;
    auto take = [&](size_t bytes) { char* q = p; p += bytes; return q; };
    c->wf.ray_o = (float4*)take(16ull * n);
    c->wf.ray_d = (float4*)take(16ull * n);
    c->wf.beta = (float4*)take(16ull * n);
    c->wf.unit = (uint4*)take(16ull * n);
    c->wf.sum = (float4*)take(16ull * n);
    c->wf.hit = (uint4*)take(16ull * n);
    c->wf.sumsq = want_sumsq ? (float4*)take(16ull * n) : nullptr;
    c->wf.queue = (uint32_t*)take(4ull * VKW_CLASSES * n);
    c->wf.qcount = (uint32_t*)take(64);
    c->wf.unit_head = (unsigned long long*)take(64);
    c->wf.n_slots = n;
    c->wf_has_sumsq = want_sumsq;
    if (!c->wf_host_counts) CU(c, cudaMallocHost((void**)&c->wf_host_counts, 64));
    return VK_OK;
}

// generate -> (extend -> shade)* until the pool has drained.  The host cannot see the pool, so it
// launches iterations in batches and reads the queue counters of the last extend after each batch
// (one 32-byte D2H + stream sync); an iteration on a drained pool is two empty launches.
static int wf_render(vk_ctx* c, bool strict, const FlatProgram* flat, const DCamera& dc, const RenderArgs& a, const RenderBuffers& b,
                     bool want_sumsq, uint32_t* launches) {
    int rc = wf_ensure(c, want_sumsq);
    if (rc != VK_OK) return rc;
    WfState w = c->wf;
    if (!want_sumsq) w.sumsq = nullptr;
    w.n_pixels = a.width * a.height;
    w.n_units = (unsigned long long)w.n_pixels * a.n_planes;
    const unsigned long long head0 = w.n_units < w.n_slots ? w.n_units : w.n_slots;
    CU(c, cudaMemsetAsync(w.qcount, 0, 64, c->stream));
    CU(c, cudaMemcpyAsync(w.unit_head, &head0, sizeof(head0), cudaMemcpyHostToDevice, c->stream));
    CU(c, strict ? vkstrict::launch_wf_generate(dc, a, w, c->stream) : vkfast::launch_wf_generate(dc, a, w, c->stream));
    ++*launches;
    // no sample ends before its first segment: at least spp_count * n_pixels / n_slots iterations
    const unsigned long long min_iters = ((unsigned long long)w.n_pixels * a.spp_count + w.n_slots - 1) / w.n_slots;
    unsigned long long next_check = min_iters > 8 ? min_iters : 8;
    const unsigned long long max_iters = (min_iters + 2) * (unsigned long long)(a.max_depth ? a.max_depth : 1) + 64; // every path is <= max_depth segments
    for (unsigned long long it = 0;; ++it) {
        if (it > max_iters) return fail(c, VK_ERR_CUDA, "wavefront: the slot pool did not drain (internal error)");
        const uint32_t set = (uint32_t)(it & 1u);
        CU(c, strict ? vkstrict::launch_wf_extend(c->scene, flat, a, w, b, set, c->stream)
                     : vkfast::launch_wf_extend(c->scene, flat, a, w, b, set, c->stream));
        CU(c, strict ? vkstrict::launch_wf_shade(c->scene, dc, a, w, b, set, c->stream)
                     : vkfast::launch_wf_shade(c->scene, dc, a, w, b, set, c->stream));
        *launches += flat ? 2 : 3; // BVH scenes: extend + classify + shade
        if (it + 1 >= next_check) {
            CU(c, cudaMemcpyAsync(c->wf_host_counts, w.qcount + set * VKW_CLASSES, VKW_CLASSES * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
            CU(c, cudaStreamSynchronize(c->stream));
            const uint32_t live = c->wf_host_counts[0] + c->wf_host_counts[1] + c->wf_host_counts[2] + c->wf_host_counts[3];
            if (live == 0) break;
            // the pool is still busy: look again after about a quarter of what is left at most, at least 8 iterations
            next_check = it + 1 + (live == w.n_slots ? 32 : 8);
        }
    }
    return VK_OK;
}


// VK_VARIANT_AUTO, by the measurements recorded in profiles/ and DESIGN.md section 4:
//   scene small enough for the flat program  -> STAGED (Cornell 600x600x1000: 67 ms against 80 ms for
//       the lane megakernel and 138 ms for the global wavefront; 20 of 32 lanes active against 15,
//       instruction-fetch stalls gone)
//   BVH scenes                               -> MEGAKERNEL (final scene: 49 ms against 78 ms staged and
//       98 ms wavefront: while-while traversal diverges whatever the staging, and the lane
//       megakernel keeps 32 warps per SM against 24)
static uint32_t choose_variant(const vk_ctx* c, const vk_render_params* P) {
    if (P->variant != VK_VARIANT_AUTO) return P->variant;
    const bool flat = c->flat.n && !(P->flags & VK_FLAG_FORCE_BVH);
    // (a hybrid program holds subtrees whose traversal lengths vary: the lane megakernel, not the
    // barrier-synchronised staged kernel, runs it)
    return flat && c->flat.n_bvh == 0 ? (uint32_t)VK_VARIANT_STAGED : (uint32_t)VK_VARIANT_MEGAKERNEL;
}

// Lane megakernel for a BVH scene: static (one whole ray per lane and loop iteration) or dynamic
// (resumable traversal under warp votes with re-fill).  Measured on B200 (profiles/r1_configs_dyn.log):
// the dynamic kernel wins where traversal lengths have a long tail -- 10^6 spheres, 66 node visits per
// ray: 62 ms against 96 ms, 15.6 against 4.5 lanes active -- and loses on the small heterogeneous
// scenes (final scene 57 ms against 44 ms, random spheres 13.8 against 13.1 ms), where shading is a
// larger share of the work and runs with few lanes during a re-fill.  VECCHIO_MEGA=static|dynamic overrides.
static bool use_dynamic_megakernel(const vk_ctx* c) {
    if (const char* e = std::getenv("VECCHIO_MEGA")) {
        if (!std::strcmp(e, "static")) return false;
        if (!std::strcmp(e, "dynamic")) return true;
    }
    return c->n_nodes >= 65536u;
}

// shared body of vk_render / vk_render_device
static int render_into(vk_ctx* c, const vk_camera* cam, const vk_render_params* P, float* d_sum, float* d_sumsq, vk_stats* stats) {
    if (!c) return VK_ERR_INVALID;
    if (!c->has_scene) return fail(c, VK_ERR_NO_SCENE, "render: no scene uploaded");
    if (!cam || !P || !d_sum) return fail(c, VK_ERR_INVALID, "render: null argument");
    if (P->width < 2 || P->height < 2) return fail(c, VK_ERR_INVALID, "render: width/height must be >= 2 ((width-1) divides, src/main.rs:187)");
    if (P->spp == 0 || P->spp_begin >= P->spp) return fail(c, VK_ERR_INVALID, "render: bad spp / spp_begin");
    const uint32_t count = P->spp_count ? P->spp_count : P->spp - P->spp_begin;
    if ((uint64_t)P->spp_begin + count > P->spp) return fail(c, VK_ERR_INVALID, "render: sample slice exceeds spp");
    if ((uint64_t)P->width * P->height > 0x7FFFFFFFull / 3) return fail(c, VK_ERR_INVALID, "render: image too large");
    if (P->variant > VK_VARIANT_STAGED) return fail(c, VK_ERR_INVALID, "render: unknown variant");
    if (!(cam->time0 < cam->time1)) return fail(c, VK_ERR_INVALID, "render: camera time0 >= time1 (gen_range panics, src/main.rs:118)");
    CU(c, cudaSetDevice(c->device));
    const bool strict = (P->flags & VK_FLAG_STRICT_MATH) != 0;

    const FlatProgram* flat = (c->flat.n && !(P->flags & VK_FLAG_FORCE_BVH)) ? &c->flat : nullptr;
    int bps = 0, bt = 0;
    const bool legacy = (P->flags & VK_FLAG_LEGACY_SCATTER) != 0;
    if (legacy && c->has_specdiffuse)
        return fail(c, VK_ERR_UNSUPPORTED, "render: SpecDiffuse has no legacy scatter (the reference's default unwraps a missing specular ray and panics, src/material.rs:21-28)");
    if (!legacy && c->scene.n_lights == 0)
        return fail(c, VK_ERR_INVALID, "render: empty light list (the reference panics: choose().unwrap(), src/hittable.rs:431); only VK_FLAG_LEGACY_SCATTER renders without lights");
    CU(c, strict ? vkstrict::megakernel_occupancy(flat != nullptr, c->scene.has_media, legacy, &bps, &bt)
                 : vkfast::megakernel_occupancy(flat != nullptr, c->scene.has_media, legacy, &bps, &bt));
    if (bps < 1) bps = 1;
    const int grid = c->sm_count * bps;
    const uint32_t resident_warps = (uint32_t)grid * (uint32_t)bt / 32u;

    RenderArgs a{};
    a.width = P->width;
    a.height = P->height;
    a.spp_begin = P->spp_begin;
    a.spp_count = count;
    a.max_depth = P->max_depth;
    a.seed_lo = (uint32_t)P->seed;
    a.seed_hi = (uint32_t)(P->seed >> 32);
    a.background = make_float3(P->background[0], P->background[1], P->background[2]);
    a.flags = P->flags;
    a.tiles_x = (P->width + 7) / 8;
    a.tiles_y = (P->height + 3) / 4;
    // Sample blocks ("units") and their planes of partial sums.  A unit's samples are traced one after
    // the other by whoever owns the unit, so a unit of k samples on a pixel whose paths run to
    // max_depth holds a lane for k * max_depth segments while the rest of the GPU has drained
    // (measured: 8-sample units cost Cornell a 2.8 ms tail, 523 iterations in the slowest CTA).
    // HBM is cheap here: one plane per SAMPLE when that fits the plane budget (default 6 GiB:
    // Cornell 600x600x1000 = 4.3 GB of planes, written once with plain stores and summed in order by
    // k_reduce_chunks at HBM speed), otherwise the smallest block that fits.  The block size depends
    // only on the call's parameters, so a render is bit-identical per (seed, spp slice, size).
    const size_t plane = (size_t)P->width * P->height * 3;
    RenderBuffers b{};
    b.counters = c->counters;
    b.debug = c->debug;
    {
        size_t budget = 6ull << 30;
        if (const char* e = std::getenv("VECCHIO_PLANE_BUDGET_MB")) {
            const long v = std::atol(e);
            if (v > 0) budget = (size_t)v << 20;
        }
        const size_t plane_bytes = plane * sizeof(float) * (d_sumsq ? 2 : 1);
        size_t max_planes = budget / plane_bytes;
        if (max_planes < 1) max_planes = 1;
        for (;;) { // a device with less free memory than the budget gets fewer, larger blocks
            a.unit_spp = (uint32_t)((count + max_planes - 1) / max_planes);
            a.n_planes = (count + a.unit_spp - 1) / a.unit_spp;
            if (a.n_planes == 1) {
                b.partial_sum = d_sum;
                b.partial_sumsq = d_sumsq;
                break;
            }
            const int rc = ensure(c, &c->partial, &c->partial_floats, plane * a.n_planes * (d_sumsq ? 2 : 1));
            if (rc == VK_OK) {
                b.partial_sum = c->partial;
                b.partial_sumsq = d_sumsq ? c->partial + plane * a.n_planes : nullptr;
                break;
            }
            if (rc != VK_ERR_OOM) return rc;
            cudaGetLastError(); // clear the allocation failure
            max_planes = a.n_planes / 2;
            if (max_planes < 1) max_planes = 1;
        }
    }
    // Chunks of whole blocks, sized for ~48 work items per resident warp (small items keep the
    // end-of-kernel tail short; an item costs one atomic).
    const uint32_t n_tiles = a.tiles_x * a.tiles_y;
    uint32_t n_chunks = (48u * resident_warps + n_tiles - 1) / n_tiles;
    if (n_chunks > a.n_planes) n_chunks = a.n_planes;
    if (n_chunks < 1) n_chunks = 1;
    const uint32_t planes_per_chunk = (a.n_planes + n_chunks - 1) / n_chunks;
    a.chunk_spp = planes_per_chunk * a.unit_spp;
    a.n_chunks = (count + a.chunk_spp - 1) / a.chunk_spp;

    CU(c, cudaMemsetAsync(c->counters + 2, 0, sizeof(unsigned long long), c->stream)); // work-queue head only
    CU(c, cudaEventRecord(c->ev0, c->stream));
    const DCamera dc = to_dcam(cam);
    uint32_t launches = 0;
    const uint32_t variant = legacy ? (uint32_t)VK_VARIANT_MEGAKERNEL : choose_variant(c, P); // legacy: lane megakernel only
    if (variant == VK_VARIANT_WAVEFRONT) {
        int rc = wf_render(c, strict, flat, dc, a, b, d_sumsq != nullptr, &launches);
        if (rc != VK_OK) return rc;
    } else if (variant == VK_VARIANT_STAGED) {
        CU(c, cudaMemsetAsync(c->counters + 5, 0xFF, 2 * sizeof(unsigned long long), c->stream)); // CTA start / first end: minima
        CU(c, cudaMemsetAsync(c->counters + 7, 0, sizeof(unsigned long long), c->stream));        // last end: maximum
        CU(c, cudaMemsetAsync(c->counters + 2, 0, sizeof(unsigned long long), c->stream)); // unit queue head
        // a "simple" scene runs the build of the same kernel with the unreachable code compiled out
        // (the staged kernel's K-ray flat trace has no subtree entries: a hybrid program means its BVH path)
        const FlatProgram* sflat = flat && flat->n_bvh == 0 ? flat : nullptr;
        const bool simple = !strict && sflat && c->simple_scene && !std::getenv("VECCHIO_NO_SIMPLE");
        CU(c, strict   ? vkstrict::launch_staged(c->scene, sflat, dc, a, b, c->counters + 2, c->sm_count, c->stream)
              : simple ? vkfast_simple::launch_staged(c->scene, sflat, dc, a, b, c->counters + 2, c->sm_count, c->stream)
                       : vkfast::launch_staged(c->scene, sflat, dc, a, b, c->counters + 2, c->sm_count, c->stream));
        launches = 1;
    } else if (!flat && !legacy && use_dynamic_megakernel(c)) {
        CU(c, cudaMemsetAsync(c->counters + 2, 0, sizeof(unsigned long long), c->stream)); // unit queue head
        CU(c, strict ? vkstrict::launch_megakernel_dyn(c->scene, dc, a, b, c->counters + 2, c->sm_count, c->stream)
                     : vkfast::launch_megakernel_dyn(c->scene, dc, a, b, c->counters + 2, c->sm_count, c->stream));
        launches = 1;
    } else {
        CU(c, strict ? vkstrict::launch_megakernel(c->scene, flat, dc, a, b, grid, legacy, c->stream)
                     : vkfast::launch_megakernel(c->scene, flat, dc, a, b, grid, legacy, c->stream));
        launches = 1;
    }
    if (a.n_planes > 1) {
        const unsigned g = (unsigned)((plane + 255) / 256);
        k_reduce_chunks<<<g, 256, 0, c->stream>>>(b.partial_sum, a.n_planes, plane, d_sum);
        ++launches;
        if (d_sumsq) {
            k_reduce_chunks<<<g, 256, 0, c->stream>>>(b.partial_sumsq, a.n_planes, plane, d_sumsq);
            ++launches;
        }
        CU(c, cudaGetLastError());
    }
    CU(c, cudaEventRecord(c->ev1, c->stream));
    c->launches += launches;
    c->paths += (uint64_t)P->width * P->height * count;
    if (stats) { // reading the counters synchronises; with stats == NULL the call is fully asynchronous
        int rc = vk_flush_stats(c, stats);
        if (rc != VK_OK) return rc;
        CU(c, cudaEventElapsedTime(&stats->ms_kernels, c->ev0, c->ev1));
        stats->ms_total = stats->ms_kernels;
        stats->variant = variant;
    }
    return VK_OK;
}

int vk_render_device(vk_ctx* c, const vk_camera* cam, const vk_render_params* P, float* d_sum, float* d_sumsq, vk_stats* stats) {
    return render_into(c, cam, P, d_sum, d_sumsq, stats);
}

int vk_finalize_device(vk_ctx* c, const float* d_sum, float* d_rgb, size_t n, uint32_t spp) {
    if (!c || !d_sum || !d_rgb || spp == 0) return fail(c, VK_ERR_INVALID, "vk_finalize_device: bad argument");
    CU(c, cudaSetDevice(c->device));
    if (n) k_finalize<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(d_sum, d_rgb, n, (float)spp);
    c->launches += n ? 1 : 0;
    CU(c, cudaGetLastError());
    return VK_OK;
}

int vk_set_stream(vk_ctx* c, void* stream) {
    if (!c) return VK_ERR_INVALID;
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaStreamSynchronize(c->stream));
    c->stream = stream ? (cudaStream_t)stream : c->own_stream;
    return VK_OK;
}

int vk_flush_stats(vk_ctx* c, vk_stats* stats) {
    if (!c || !stats) return VK_ERR_INVALID;
    CU(c, cudaSetDevice(c->device));
    unsigned long long h[5] = {0, 0, 0, 0, 0};
    CU(c, cudaMemcpyAsync(h, c->counters, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaMemsetAsync(c->counters, 0, 2 * sizeof(unsigned long long), c->stream));
    CU(c, cudaMemsetAsync(c->counters + 3, 0, 2 * sizeof(unsigned long long), c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    std::memset(stats, 0, sizeof(*stats));
    stats->paths = c->paths;
    stats->rays = h[0];
    stats->dropped_samples = h[1];
    stats->node_visits = h[3];
    stats->prim_tests = h[4];
    stats->launches = (uint32_t)c->launches;
    stats->variant = VK_VARIANT_MEGAKERNEL;
    c->paths = 0;
    c->launches = 0;
    return VK_OK;
}

int vk_render(vk_ctx* c, const vk_camera* cam, const vk_render_params* P, float* out_rgb, float* out_sumsq, vk_stats* stats) {
    if (!c) return VK_ERR_INVALID;
    if (!P || !out_rgb) return fail(c, VK_ERR_INVALID, "vk_render: null argument");
    const size_t plane = (size_t)P->width * P->height * 3;
    CU(c, cudaSetDevice(c->device));
    int rc = ensure(c, &c->frame, &c->frame_floats, plane * 3);
    if (rc != VK_OK) return rc;
    if (c->pinned_floats < plane * 2) {
        if (c->pinned) cudaFreeHost(c->pinned);
        c->pinned = nullptr;
        c->pinned_floats = 0;
        CU(c, cudaMallocHost((void**)&c->pinned, plane * 2 * sizeof(float)));
        c->pinned_floats = plane * 2;
    }
    float *d_sum = c->frame, *d_sq = out_sumsq ? c->frame + plane : nullptr, *d_rgb = c->frame + 2 * plane;
    CU(c, cudaEventRecord(c->ev2, c->stream));
    vk_stats st{};
    rc = render_into(c, cam, P, d_sum, d_sq, &st);
    if (rc != VK_OK) return rc;
    k_finalize<<<(unsigned)((plane + 255) / 256), 256, 0, c->stream>>>(d_sum, d_rgb, plane, (float)P->spp);
    CU(c, cudaGetLastError());
    CU(c, cudaMemcpyAsync(c->pinned, d_rgb, plane * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    if (out_sumsq) CU(c, cudaMemcpyAsync(c->pinned + plane, d_sq, plane * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaEventRecord(c->ev1, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    std::memcpy(out_rgb, c->pinned, plane * sizeof(float));
    if (out_sumsq) std::memcpy(out_sumsq, c->pinned + plane, plane * sizeof(float));
    st.launches += 1;
    c->launches = 0;
    CU(c, cudaEventElapsedTime(&st.ms_total, c->ev2, c->ev1));
    if (stats) *stats = st;
    return VK_OK;
}

int vk_render_rgb8(vk_ctx* c, const vk_camera* cam, const vk_render_params* P, uint8_t* out_rgb8, vk_stats* stats) {
    if (!c) return VK_ERR_INVALID;
    if (!P || !out_rgb8) return fail(c, VK_ERR_INVALID, "vk_render_rgb8: null argument");
    const size_t plane = (size_t)P->width * P->height * 3;
    CU(c, cudaSetDevice(c->device));
    int rc = ensure(c, &c->frame, &c->frame_floats, plane * 3);
    if (rc != VK_OK) return rc;
    if (c->pinned_floats < plane * 2) {
        if (c->pinned) cudaFreeHost(c->pinned);
        c->pinned = nullptr;
        c->pinned_floats = 0;
        CU(c, cudaMallocHost((void**)&c->pinned, plane * 2 * sizeof(float)));
        c->pinned_floats = plane * 2;
    }
    float* d_sum = c->frame;
    uint8_t* d_rgb8 = (uint8_t*)(c->frame + 2 * plane);
    CU(c, cudaEventRecord(c->ev2, c->stream));
    vk_stats st{};
    rc = render_into(c, cam, P, d_sum, nullptr, &st);
    if (rc != VK_OK) return rc;
    k_to_color<<<(unsigned)((plane + 255) / 256), 256, 0, c->stream>>>(d_sum, d_rgb8, P->width, P->height, (float)P->spp);
    CU(c, cudaGetLastError());
    CU(c, cudaMemcpyAsync(c->pinned, d_rgb8, plane, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaEventRecord(c->ev1, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    std::memcpy(out_rgb8, c->pinned, plane);
    st.launches += 1;
    c->launches = 0;
    CU(c, cudaEventElapsedTime(&st.ms_total, c->ev2, c->ev1));
    if (stats) *stats = st;
    return VK_OK;
}

int vk_intersect(vk_ctx* c, const vk_ray* rays, size_t n, const float* medium_xi, uint32_t flags, vk_hit* out) {
    if (!c) return VK_ERR_INVALID;
    if (!c->has_scene) return fail(c, VK_ERR_NO_SCENE, "vk_intersect: no scene uploaded");
    if (n == 0) return VK_OK;
    if (!rays || !out) return fail(c, VK_ERR_INVALID, "vk_intersect: null argument");
    CU(c, cudaSetDevice(c->device));
    vk_ray* d_rays = nullptr;
    vk_hit* d_hits = nullptr;
    float* d_xi = nullptr;
    int rc = VK_OK;
    cudaError_t e;
#define STEP(call)                                                                                                     \
    if (rc == VK_OK && (e = (call)) != cudaSuccess) rc = fail(c, VK_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e));
    STEP(cudaMalloc((void**)&d_rays, n * sizeof(vk_ray)))
    STEP(cudaMalloc((void**)&d_hits, n * sizeof(vk_hit)))
    if (medium_xi) {
        STEP(cudaMalloc((void**)&d_xi, n * VK_MEDIUM_XI_SLOTS * sizeof(float)))
        STEP(cudaMemcpyAsync(d_xi, medium_xi, n * VK_MEDIUM_XI_SLOTS * sizeof(float), cudaMemcpyHostToDevice, c->stream))
    }
    STEP(cudaMemcpyAsync(d_rays, rays, n * sizeof(vk_ray), cudaMemcpyHostToDevice, c->stream))
    const FlatProgram* flat = (c->flat.n && !(flags & VK_FLAG_FORCE_BVH)) ? &c->flat : nullptr;
    STEP((flags & VK_FLAG_STRICT_MATH) ? vkstrict::launch_intersect(c->scene, flat, d_rays, n, d_xi, d_hits, c->stream)
                                       : vkfast::launch_intersect(c->scene, flat, d_rays, n, d_xi, d_hits, c->stream))
    STEP(cudaMemcpyAsync(out, d_hits, n * sizeof(vk_hit), cudaMemcpyDeviceToHost, c->stream))
    STEP(cudaStreamSynchronize(c->stream))
#undef STEP
    cudaFree(d_rays);
    cudaFree(d_hits);
    cudaFree(d_xi);
    return rc;
}

int vk_measure_peaks(vk_ctx* c, float* fp32_tflops, float* l2_gbs) {
    if (!c) return VK_ERR_INVALID;
    CU(c, cudaSetDevice(c->device));
    float* out = nullptr;
    const int blocks = c->sm_count * 8, threads = 256, iters = 4096;
    CU(c, cudaMalloc((void**)&out, (size_t)blocks * threads * sizeof(float)));
    float best_ms = 1e30f, ms = 0.f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(c->ev0, c->stream);
        k_ffma_peak<<<blocks, threads, 0, c->stream>>>(out, iters);
        cudaEventRecord(c->ev1, c->stream);
        cudaStreamSynchronize(c->stream);
        cudaEventElapsedTime(&ms, c->ev0, c->ev1);
        if (r > 0 && ms < best_ms) best_ms = ms;
    }
    if (fp32_tflops) *fp32_tflops = (float)((double)blocks * threads * iters * 16.0 * 8.0 * 2.0 / (best_ms * 1e-3) / 1e12);
    cudaFree(out);
    const size_t bytes = 32ull << 20; // 32 MiB: well inside the 126 MB L2
    float4* buf = nullptr;
    float* sink = nullptr;
    CU(c, cudaMalloc((void**)&buf, bytes));
    CU(c, cudaMalloc((void**)&sink, sizeof(float)));
    CU(c, cudaMemsetAsync(buf, 0, bytes, c->stream));
    const int reps = 20;
    best_ms = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(c->ev0, c->stream);
        k_l2_read<<<c->sm_count * 8, 256, 0, c->stream>>>(buf, bytes / sizeof(float4), reps, sink);
        cudaEventRecord(c->ev1, c->stream);
        cudaStreamSynchronize(c->stream);
        cudaEventElapsedTime(&ms, c->ev0, c->ev1);
        if (r > 0 && ms < best_ms) best_ms = ms;
    }
    if (l2_gbs) *l2_gbs = (float)((double)bytes * reps / (best_ms * 1e-3) / 1e9);
    cudaFree(buf);
    cudaFree(sink);
    CU(c, cudaGetLastError());
    return VK_OK;
}

// Debug hook: the raw counter block (see RenderBuffers); [5..7] = first CTA start, first and last CTA end of the
// last staged launch (globaltimer ns).
int vk_debug_counters(vk_ctx* c, unsigned long long out[8]) {
    if (!c || !out) return VK_ERR_INVALID;
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaStreamSynchronize(c->stream));
    CU(c, cudaMemcpy(out, c->counters, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    return VK_OK;
}
int vk_debug_ctas(vk_ctx* c, unsigned long long* out, size_t n_ctas) {
    if (!c || !out || n_ctas > VK_DEBUG_CTAS) return VK_ERR_INVALID;
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaStreamSynchronize(c->stream));
    CU(c, cudaMemcpy(out, c->debug, n_ctas * 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    return VK_OK;
}

// Test hook: Philox4x32-10 on the device for the Random123 known-answer vectors.
int vk_selftest_philox(vk_ctx* c, const uint32_t ctr_key6[6], uint32_t out4[4]) {
    if (!c || !ctr_key6 || !out4) return VK_ERR_INVALID;
    CU(c, cudaSetDevice(c->device));
    uint32_t* d = nullptr;
    CU(c, cudaMalloc((void**)&d, 10 * sizeof(uint32_t)));
    cudaMemcpyAsync(d, ctr_key6, 6 * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream);
    cudaError_t e = vkfast::launch_philox_kat(d, d + 6, c->stream);
    cudaMemcpyAsync(out4, d + 6, 4 * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream);
    cudaStreamSynchronize(c->stream);
    cudaFree(d);
    CU(c, e);
    CU(c, cudaGetLastError());
    return VK_OK;
}

} // extern "C"
