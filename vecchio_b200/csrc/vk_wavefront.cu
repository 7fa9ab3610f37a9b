// vk_wavefront.cu -- the wavefront variant of the sample loop (see WfState in vk_internal.h).
// Compiled twice like vk_kernels.cu (namespaces vkfast / vkstrict).
//
//   k_wf_generate   every slot of the pool takes a unit and the camera ray of its first sample
//   k_wf_extend     BVHNode::hit / flat program for every live slot  (src/main.rs:130)
//                   + classification by shading class and queue append
//   k_wf_shade      material scatter + mixture-PDF sampling (src/main.rs:131-149), class by class;
//                   sample end: NaN filter, accumulate (main.rs:191-194), regenerate
#include "vk_device.cuh"

namespace VK_NS {

#define VKW_BLOCK 256

VKD uint32_t f2u(float f) { return __float_as_uint(f); }
VKD float u2f(uint32_t u) { return __uint_as_float(u); }

// Start sample `s` of the slot's unit: Camera::get_ray with depth 1 (src/main.rs:187-190).
VKD void wf_begin_sample(const DCamera& cam, const RenderArgs& a, const WfState& w, uint32_t slot, uint32_t pixel, uint32_t s) {
    PathRng rng;
    rng.pixel = pixel;
    rng.sample = s;
    rng.key = make_uint2(a.seed_lo, a.seed_hi);
    float3 o, d;
    float time;
    camera_get_ray(cam, rng, pixel % a.width, pixel / a.width, a.width, a.height, o, d, time);
    w.ray_o[slot] = make_float4(o.x, o.y, o.z, time);
    w.ray_d[slot] = make_float4(d.x, d.y, d.z, u2f(1u));           // ray_color(ray, .., 1)
    w.beta[slot] = make_float4(1.0f, 1.0f, 1.0f, u2f(s));
}
// Unit u = (sample b, pixel): b = u / n_pixels.  Pixels run row-major, so the 32 slots a
// warp fills together get 32 neighbouring pixels of a row.
VKD void wf_start_unit(const DCamera& cam, const RenderArgs& a, const WfState& w, uint32_t slot, unsigned long long u) {
    const uint32_t b = (uint32_t)(u / w.n_pixels), pixel = (uint32_t)(u - (unsigned long long)b * w.n_pixels);
    w.unit[slot] = make_uint4(pixel, 0u, 0u, 0u);
    wf_begin_sample(cam, a, w, slot, pixel, a.spp_begin + b);
}

__global__ void __launch_bounds__(VKW_BLOCK) k_wf_generate(const DCamera cam, const RenderArgs a, const WfState w) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= w.n_slots) return;
    if ((unsigned long long)i < w.n_units) wf_start_unit(cam, a, w, i, i); // unit_head starts at min(n_slots, n_units)
    else w.ray_d[i] = make_float4(0.0f, 0.0f, 0.0f, u2f(0u));
}

VKD uint32_t prim_material(const DScene& sc, uint32_t prim) {
    const uint32_t i = VKD_INDEX(prim);
    switch (VKD_TYPE(prim)) {
    case VK_T_SPHERE: return __ldg(&sc.sphere_mat[i]);
    case VK_T_MSPHERE: return f2u(__ldg(&sc.mspheres[3 * i + 2]).y);
    case VK_T_RECT: return f2u(__ldg(&sc.rects[2 * i + 1]).z);
    case VK_T_BOX: return f2u(__ldg(&sc.boxes[2 * i]).w);
    case VK_T_MEDIUM: return f2u(__ldg(&sc.media[i]).z);
    default: return 0u;
    }
}

template <bool FLAT, bool MEDIA>
VKD void wf_extend_body(const DScene& sc, const FlatProgram* flat, const RenderArgs& a, const WfState& w, const RenderBuffers& buf,
                        uint32_t set) {
    __shared__ uint32_t s_cnt[VKW_CLASSES], s_base[VKW_CLASSES], s_stat[2];
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31u;
    if (threadIdx.x < VKW_CLASSES) s_cnt[threadIdx.x] = 0u;
    if (threadIdx.x < 2) s_stat[threadIdx.x] = 0u;
    __syncthreads();
    bool live = false;
    uint32_t cls = 0;
    TraceCounters tc = {0u, 0u};
    if (i < w.n_slots) {
        const float4 rd = w.ray_d[i];
        const uint32_t depth = f2u(rd.w);
        if (depth) {
            live = true;
            const float4 ro = w.ray_o[i];
            MediumXi xi;
            xi.table = nullptr;
            xi.depth = depth;
            xi.rng.key = make_uint2(a.seed_lo, a.seed_hi);
            xi.rng.pixel = 0;
            xi.rng.sample = 0;
            if (MEDIA) {
                xi.rng.pixel = w.unit[i].x;
                xi.rng.sample = f2u(w.beta[i].w);
            }
            const TraceHit h = FLAT ? trace_flat<MEDIA>(sc, *flat, f3(ro), f3(rd), ro.w, 0.001f, CUDART_INF_F, xi, tc)
                                    : trace<MEDIA>(sc, f3(ro), f3(rd), ro.w, 0.001f, CUDART_INF_F, xi, tc); // src/main.rs:130
            w.hit[i] = make_uint4(f2u(h.t), h.prim, h.inst, h.face);
            if (h.prim != VK_REF_NONE) {
                const uint32_t mtype = __ldg(&sc.materials[prim_material(sc, h.prim)]).x;
                cls = mtype == VK_M_DIFFUSE_LIGHT ? VKW_TERMINATE : mtype == VK_M_DIELECTRIC ? VKW_DIELECTRIC : mtype == VK_M_METAL ? VKW_METAL : VKW_DIFFUSE;
            }
        }
    }
    // queue append: ballot per class, one shared-memory atomic per warp, one global atomic per block
    uint32_t my_rank = 0;
#pragma unroll
    for (uint32_t c = 0; c < VKW_CLASSES; ++c) {
        const uint32_t m = __ballot_sync(0xFFFFFFFFu, live && cls == c);
        if (m) {
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(&s_cnt[c], (uint32_t)__popc(m));
            base = __shfl_sync(0xFFFFFFFFu, base, 0);
            if (live && cls == c) my_rank = base + __popc(m & ((1u << lane) - 1u));
        }
    }
    uint32_t wn = tc.nodes, wp = tc.prims;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        wn += __shfl_xor_sync(0xFFFFFFFFu, wn, off);
        wp += __shfl_xor_sync(0xFFFFFFFFu, wp, off);
    }
    if (lane == 0 && (wn | wp)) {
        atomicAdd(&s_stat[0], wn);
        atomicAdd(&s_stat[1], wp);
    }
    __syncthreads();
    if (threadIdx.x < VKW_CLASSES) s_base[threadIdx.x] = s_cnt[threadIdx.x] ? atomicAdd(&w.qcount[set * VKW_CLASSES + threadIdx.x], s_cnt[threadIdx.x]) : 0u;
    if (threadIdx.x == VKW_CLASSES && s_stat[0]) atomicAdd(&buf.counters[3], (unsigned long long)s_stat[0]);
    if (threadIdx.x == VKW_CLASSES + 1 && s_stat[1]) atomicAdd(&buf.counters[4], (unsigned long long)s_stat[1]);
    __syncthreads();
    if (live) w.queue[(size_t)cls * w.n_slots + s_base[cls] + my_rank] = i;
}

template <bool MEDIA>
__global__ void __launch_bounds__(VKW_BLOCK) k_wf_extend(const DScene sc, const RenderArgs a, const WfState w, const RenderBuffers buf, uint32_t set) {
    wf_extend_body<false, MEDIA>(sc, nullptr, a, w, buf, set);
}
template <bool MEDIA>
__global__ void __launch_bounds__(VKW_BLOCK) k_wf_extend_flat(const DScene sc, const __grid_constant__ FlatProgram flat, const RenderArgs a,
                                                              const WfState w, const RenderBuffers buf, uint32_t set) {
    wf_extend_body<true, MEDIA>(sc, &flat, a, w, buf, set);
}

__global__ void __launch_bounds__(VKW_BLOCK) k_wf_shade(const DScene sc, const DCamera cam, const RenderArgs a, const WfState w,
                                                        const RenderBuffers buf, uint32_t set) {
    __shared__ uint32_t s_need, s_drop;
    __shared__ unsigned long long s_ubase;
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31u;
    if (threadIdx.x == 0) {
        s_need = 0u;
        s_drop = 0u;
    }
    const uint32_t c0 = w.qcount[set * VKW_CLASSES + 0], c1 = w.qcount[set * VKW_CLASSES + 1], c2 = w.qcount[set * VKW_CLASSES + 2],
                   c3 = w.qcount[set * VKW_CLASSES + 3];
    const uint32_t total = c0 + c1 + c2 + c3;
    if (j == 0) { // the next iteration's counters; this iteration's ray count
#pragma unroll
        for (uint32_t c = 0; c < VKW_CLASSES; ++c) w.qcount[(set ^ 1u) * VKW_CLASSES + c] = 0u;
        w.qcount[2 * VKW_CLASSES + (set ^ 1u)] = 0u; // slot head of the dynamic extend
        atomicAdd(&buf.counters[0], (unsigned long long)total);
    }
    __syncthreads();
    if ((uint32_t)(blockIdx.x * blockDim.x) >= total) return; // whole block beyond the queues
    const bool in = j < total;
    bool need_unit = false, dropped = false;
    uint32_t slot = 0;
    if (in) {
        const uint32_t cls = j < c0 ? 0u : (j < c0 + c1 ? 1u : (j < c0 + c1 + c2 ? 2u : 3u));
        const uint32_t off = cls == 0 ? 0u : (cls == 1 ? c0 : (cls == 2 ? c0 + c1 : c0 + c1 + c2));
        slot = w.queue[(size_t)cls * w.n_slots + (j - off)];
        const uint4 hq = w.hit[slot];
        const float4 ro = w.ray_o[slot], rd = w.ray_d[slot], bt = w.beta[slot];
        const uint4 un = w.unit[slot];
        float3 o = f3(ro), d = f3(rd), beta = f3(bt), L = f3(0.0f, 0.0f, 0.0f);
        float time = ro.w;
        uint32_t depth = f2u(rd.w);
        const uint32_t sample = f2u(bt.w);
        bool alive, valid = true;
        if (hq.y == VK_REF_NONE) {
            L = beta * miss_color(a, d); // src/main.rs:151
            alive = false;
        } else {
            PathRng rng;
            rng.pixel = un.x;
            rng.sample = sample;
            rng.key = make_uint2(a.seed_lo, a.seed_hi);
            TraceHit h;
            h.t = u2f(hq.x);
            h.prim = hq.y;
            h.inst = hq.z;
            h.face = hq.w & 0xFFu; // (the dynamic extend keeps the shading class in the upper bits)
            HitRecD rec;
            resolve_hit(sc, h, o, d, time, false, rec);
            alive = shade(sc, rec, rng, depth, o, d, time, beta, L, valid);
            if (alive && ++depth > a.max_depth) alive = false; // `depth > MAX_DEPTH` -> 0 (src/main.rs:126)
            if (alive && !(finite3(d) && finite3(o))) {          // see the megakernel: the reference's sample is NaN
                valid = false;
                alive = false;
            }
        }
        if (alive) {
            w.ray_o[slot] = make_float4(o.x, o.y, o.z, time);
            w.ray_d[slot] = make_float4(d.x, d.y, d.z, u2f(depth));
            w.beta[slot] = make_float4(beta.x, beta.y, beta.z, bt.w);
        } else { // sample finished: NaN/Inf filter of src/main.rs:191-194, then the slot's next unit
            const bool keep = valid && finite3(L);
            dropped = !keep;
            if (keep) accumulate_sample(buf, un.x, L);
            need_unit = true;
        }
    }
    // next units: ballot + shared-memory atomic per warp, one global atomic per block
    const uint32_t m = __ballot_sync(0xFFFFFFFFu, need_unit);
    const uint32_t md = __ballot_sync(0xFFFFFFFFu, dropped);
    uint32_t rank = 0;
    if (m) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(&s_need, (uint32_t)__popc(m));
        rank = __shfl_sync(0xFFFFFFFFu, base, 0) + __popc(m & ((1u << lane) - 1u));
    }
    if (md && lane == 0) atomicAdd(&s_drop, (uint32_t)__popc(md));
    __syncthreads();
    if (threadIdx.x == 0) {
        s_ubase = s_need ? atomicAdd(w.unit_head, (unsigned long long)s_need) : 0ull;
        if (s_drop) atomicAdd(&buf.counters[1], (unsigned long long)s_drop);
    }
    __syncthreads();
    if (need_unit) {
        const unsigned long long u = s_ubase + rank;
        if (u < w.n_units) wf_start_unit(cam, a, w, slot, u);
        else w.ray_d[slot] = make_float4(0.0f, 0.0f, 0.0f, u2f(0u)); // pool drains: slot idle from now on
    }
}

// ---------------------------------------------------------------------------------------------------
// extend for BVH scenes: persistent warps with dynamic fetch.  Traversal lengths have a long tail
// (10^6 spheres: mean 28 four-wide visits, some rays hundreds), so a thread is not tied to a slot:
// the warp runs the resumable traversal (Trav) in bounded while-while rounds and a lane whose ray is
// finished takes the next untraced slot from a global counter.  With 2^19 rays per launch the tail
// is paid once per launch, not once per warp.  The hit is stored with its shading class;
// k_wf_classify then builds the class queues (block-aggregated atomics, as in wf_extend_body).
// ---------------------------------------------------------------------------------------------------
#ifndef VKW_FETCH_MIN_IDLE
#define VKW_FETCH_MIN_IDLE 8u
#endif
#ifndef VKW_NODE_STEPS
#define VKW_NODE_STEPS 4
#endif
template <bool MEDIA>
__global__ void __launch_bounds__(128, 6) k_wf_extend_dyn(const DScene sc, const RenderArgs a, const WfState w, const RenderBuffers buf,
                                                          uint32_t set) {
    const uint32_t lane = threadIdx.x & 31u, lanes_below = (1u << lane) - 1u;
    uint32_t* slot_head = &w.qcount[2 * VKW_CLASSES + set];
    Trav T;
    T.ref = VKD_DONE;
    T.sp = 0;
    T.enter = false;
    T.cur_inst = 0;
    T.co = T.cd = T.cinv = f3(0.0f, 0.0f, 0.0f);
    T.best.t = 0.0f;
    T.best.prim = VK_REF_NONE;
    T.best.inst = 0;
    T.best.face = 0;
    float3 o = f3(0.0f, 0.0f, 0.0f), d = o;
    float tm = 0.0f;
    uint32_t cur = 0xFFFFFFFFu;
    bool pool_empty = false;
    MediumXi xi;
    xi.table = nullptr;
    xi.depth = 0;
    xi.rng.key = make_uint2(a.seed_lo, a.seed_hi);
    xi.rng.pixel = 0;
    xi.rng.sample = 0;
    TraceCounters tc = {0u, 0u};
#pragma unroll 1
    for (;;) {
        const uint32_t m_idle = __ballot_sync(0xFFFFFFFFu, T.ref == VKD_DONE);
        const uint32_t m_need = __ballot_sync(0xFFFFFFFFu, T.ref == VKD_DONE && !pool_empty);
        if (m_need != 0u && ((uint32_t)__popc(m_idle) >= VKW_FETCH_MIN_IDLE)) {
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(slot_head, (uint32_t)__popc(m_need));
            base = __shfl_sync(0xFFFFFFFFu, base, 0);
            if (T.ref == VKD_DONE && !pool_empty) {
                const uint32_t my = base + __popc(m_need & lanes_below);
                if (my >= w.n_slots) pool_empty = true;
                else {
                    const float4 rd = w.ray_d[my];
                    if (f2u(rd.w) != 0u) {
                        const float4 ro = w.ray_o[my];
                        o = f3(ro);
                        d = f3(rd);
                        tm = ro.w;
                        xi.depth = f2u(rd.w);
                        if (MEDIA) {
                            xi.rng.pixel = w.unit[my].x;
                            xi.rng.sample = f2u(w.beta[my].w);
                        }
                        trav_init(T, sc, o, d, CUDART_INF_F); // world.hit(&r, 0.001, inf) src/main.rs:130
                        cur = my;
                    }
                }
            }
        }
        if (__ballot_sync(0xFFFFFFFFu, T.ref != VKD_DONE) == 0u) {
            if (__ballot_sync(0xFFFFFFFFu, !pool_empty) == 0u) break;
            continue; // every lane drew an idle slot (only while the pool drains): draw again
        }
#pragma unroll 1
        for (int k = 0; k < VKW_NODE_STEPS && trav_at_node(T); ++k) trav_node_step(T, sc, 0.001f, tc);
        if (T.ref != VKD_DONE && !trav_at_node(T)) trav_prim_step<MEDIA>(T, sc, o, d, tm, 0.001f, xi, tc);
        if (cur != 0xFFFFFFFFu && T.ref == VKD_DONE) { // finished: the hit and its shading class
            uint32_t cls = VKW_TERMINATE;
            if (T.best.prim != VK_REF_NONE) {
                const uint32_t mtype = __ldg(&sc.materials[prim_material(sc, T.best.prim)]).x;
                cls = mtype == VK_M_DIFFUSE_LIGHT ? VKW_TERMINATE : mtype == VK_M_DIELECTRIC ? VKW_DIELECTRIC : mtype == VK_M_METAL ? VKW_METAL : VKW_DIFFUSE;
            }
            w.hit[cur] = make_uint4(f2u(T.best.t), T.best.prim, T.best.inst, T.best.face | (cls << 8));
            cur = 0xFFFFFFFFu;
        }
    }
    uint32_t wn = tc.nodes, wp = tc.prims;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        wn += __shfl_xor_sync(0xFFFFFFFFu, wn, off);
        wp += __shfl_xor_sync(0xFFFFFFFFu, wp, off);
    }
    if (lane == 0 && (wn | wp)) {
        atomicAdd(&buf.counters[3], (unsigned long long)wn);
        atomicAdd(&buf.counters[4], (unsigned long long)wp);
    }
}
// class queues from the hits the dynamic extend stored (slot == thread)
__global__ void __launch_bounds__(VKW_BLOCK) k_wf_classify(const WfState w, uint32_t set) {
    __shared__ uint32_t s_cnt[VKW_CLASSES], s_base[VKW_CLASSES];
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31u;
    if (threadIdx.x < VKW_CLASSES) s_cnt[threadIdx.x] = 0u;
    __syncthreads();
    const bool live = i < w.n_slots && f2u(w.ray_d[i].w) != 0u;
    const uint32_t cls = live ? ((w.hit[i].w >> 8) & 3u) : 0u;
    uint32_t my_rank = 0;
#pragma unroll
    for (uint32_t c = 0; c < VKW_CLASSES; ++c) {
        const uint32_t m = __ballot_sync(0xFFFFFFFFu, live && cls == c);
        if (m) {
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(&s_cnt[c], (uint32_t)__popc(m));
            base = __shfl_sync(0xFFFFFFFFu, base, 0);
            if (live && cls == c) my_rank = base + __popc(m & ((1u << lane) - 1u));
        }
    }
    __syncthreads();
    if (threadIdx.x < VKW_CLASSES) s_base[threadIdx.x] = s_cnt[threadIdx.x] ? atomicAdd(&w.qcount[set * VKW_CLASSES + threadIdx.x], s_cnt[threadIdx.x]) : 0u;
    __syncthreads();
    if (live) w.queue[(size_t)cls * w.n_slots + s_base[cls] + my_rank] = i;
}

cudaError_t launch_wf_generate(const DCamera& cam, const RenderArgs& a, const WfState& w, cudaStream_t st) {
    k_wf_generate<<<(w.n_slots + VKW_BLOCK - 1) / VKW_BLOCK, VKW_BLOCK, 0, st>>>(cam, a, w);
    return cudaGetLastError();
}
cudaError_t launch_wf_extend(const DScene& sc, const FlatProgram* flat, const RenderArgs& a, const WfState& w, const RenderBuffers& b,
                             uint32_t set, cudaStream_t st) {
    const unsigned grid = (w.n_slots + VKW_BLOCK - 1) / VKW_BLOCK;
    if (flat && flat->n) {
        if (sc.has_media) k_wf_extend_flat<true><<<grid, VKW_BLOCK, 0, st>>>(sc, *flat, a, w, b, set);
        else k_wf_extend_flat<false><<<grid, VKW_BLOCK, 0, st>>>(sc, *flat, a, w, b, set);
    } else { // BVH scene: persistent warps with dynamic fetch, then the class queues
        int bps = 0;
        cudaError_t e = sc.has_media ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k_wf_extend_dyn<true>, 128, 0)
                                     : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k_wf_extend_dyn<false>, 128, 0);
        if (e != cudaSuccess) return e;
        int dev = 0, sms = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const unsigned pgrid = (unsigned)(sms * (bps < 1 ? 1 : bps));
        if (sc.has_media) k_wf_extend_dyn<true><<<pgrid, 128, 0, st>>>(sc, a, w, b, set);
        else k_wf_extend_dyn<false><<<pgrid, 128, 0, st>>>(sc, a, w, b, set);
        k_wf_classify<<<grid, VKW_BLOCK, 0, st>>>(w, set);
    }
    return cudaGetLastError();
}
cudaError_t launch_wf_shade(const DScene& sc, const DCamera& cam, const RenderArgs& a, const WfState& w, const RenderBuffers& b,
                            uint32_t set, cudaStream_t st) {
    k_wf_shade<<<(w.n_slots + VKW_BLOCK - 1) / VKW_BLOCK, VKW_BLOCK, 0, st>>>(sc, cam, a, w, b, set);
    return cudaGetLastError();
}

} // namespace VK_NS
