// vk_staged.cu -- the persistent megakernel, staged: a wavefront inside each CTA.
//
// Measured on B200 (profiles/r1_cornell_megakernel_*): the one-lane-one-path megakernel issues with
// 15 of 32 lanes active and stalls mostly on instruction fetch ("no_instruction"): its lanes sit in
// different materials / regeneration / traversal at the same time, so every warp walks a 70 KB body.
// The global-memory wavefront fixes the divergence but pays launch boundaries and queue traffic.
// This kernel keeps the wavefront's stages and sorts, but inside one persistent CTA:
//
//   * a pool of VKS_N path slots per CTA lives in shared memory (four 16-byte records per slot, 3 CTAs per SM);
//   * each iteration runs the stages back to back, separated by __syncthreads():
//       extend      world.hit() for every live slot, slot == thread mapping   src/main.rs:130
//       sort        the live slots are appended to the list of their shading class (warp ballot +
//                   one shared-memory atomic per warp and class)
//       shade       threads walk the class lists end to end, so a warp runs ONE material
//                   (src/main.rs:131-149); sample end: NaN filter and accumulate (:191-194), then
//                   the next sample's camera ray in place (:187-190) -- the emitter/miss class
//                   does this with full warps
//   * a slot owns a (pixel, sample) unit; the finished sample goes into the pixel's integer accumulators,
//     so the image is bit-identical to the other variants for a seed.
#include "vk_device.cuh"

namespace VK_NS {

#define VKS_T 256
#ifndef VKS_N
#define VKS_N 1024
#endif
#ifndef VKS_MINB
#define VKS_MINB 3
#endif
#define VKS_ROUNDS (VKS_N / VKS_T)
#define VKS_NEWUNIT 0x80000000u
enum { VKS_C_TERMINATE = 0, VKS_C_DIELECTRIC = 1, VKS_C_METAL = 2, VKS_C_DIFFUSE = 3, VKS_CLASSES = 4, VKS_C_IDLE = 7 };

#define VKS_CHUNK 256u // units a CTA takes from the global queue at a time
#define VKS_RING 8u    // chunk bases kept: RING * CHUNK >= N + 2 * CHUNK
#ifndef VKS_K
#define VKS_K 2 // rays a thread traces together through the flat program
#endif
struct StagedShared {
    float4 ro[VKS_N];                  // origin.xyz, time
    float4 rd[VKS_N];                  // direction.xyz, bits: depth of the segment to trace (0 = idle slot)
    float4 bt[VKS_N];                  // path weight.xyz, bits: global sample index
    uint4 hp[VKS_N];                   // hit: t bits, primitive, instance index | face << 28 | has-instance << 31; pixel
    uint16_t list[VKS_CLASSES][VKS_N]; // live slots by shading class (each class can hold the whole pool)
    uint32_t cnt[2][VKS_CLASSES];      // class counts, double buffered by iteration parity
    uint16_t late[VKS_N];              // slots whose regeneration was queued (| 0x8000: needs a new unit)
    uint32_t n_late;
    uint32_t no_units;                 // set once a slot drew a unit past the end: the global queue is empty
    uint32_t next_unit;                // CTA-local unit counter (local index n)
    uint32_t fetched;                  // local indices [0, fetched) are backed by a chunk
    uint32_t chunk_b[VKS_RING], chunk_y[VKS_RING], chunk_x[VKS_RING]; // first unit of each chunk of the ring: sample block, row, column
};

VKD uint32_t staged_class_of(const DScene& sc, uint32_t prim) {
    uint32_t mat;
    const uint32_t i = VKD_INDEX(prim);
    switch (VKD_TYPE(prim)) {
    case VK_T_SPHERE: mat = __ldg(&sc.sphere_mat[i]); break;
    case VK_T_MSPHERE: mat = __float_as_uint(__ldg(&sc.mspheres[3 * i + 2]).y); break;
    case VK_T_RECT: mat = __float_as_uint(__ldg(&sc.rects[2 * i + 1]).z); break;
    case VK_T_BOX: mat = __float_as_uint(__ldg(&sc.boxes[2 * i]).w); break;
    default: mat = __float_as_uint(__ldg(&sc.media[i]).z); break;
    }
    const uint32_t t = __ldg(&sc.materials[mat]).x;
    return t == VK_M_DIFFUSE_LIGHT ? VKS_C_TERMINATE : t == VK_M_DIELECTRIC ? VKS_C_DIELECTRIC : t == VK_M_METAL ? VKS_C_METAL : VKS_C_DIFFUSE;
}

// Units (unit = (pixel, sample block), pixels row-major) come from one global queue in chunks of
// VKS_CHUNK: thread 0 refills the CTA's chunk ring during the extend stage whenever fewer than a
// pool's worth of units is backed (one global atomic per chunk, off everybody's critical path: the
// barrier between extend and shade publishes it); inside the CTA the slots draw local indices from
// `next_unit` with a shared-memory atomic.  Fast SMs simply take more chunks, so all CTAs finish
// within a chunk of each other whatever their speed.
struct StagedCtx {
    const DCamera& cam;
    const RenderArgs& a;
    StagedShared& S;
    uint32_t n_pixels;
    unsigned long long n_units;
    unsigned long long* unit_head;
};
// one thread, every ~20 iterations: out of line, arguments by value.  The chunk's first unit is split into
// (sample block, row, column) here, once per chunk, so that the per-path code needs no division.
static __device__ __noinline__ void staged_refill(StagedShared* S, unsigned long long* unit_head, uint32_t width, uint32_t n_pixels) {
    while (S->fetched < S->next_unit + VKS_N) {
        const unsigned long long u = atomicAdd(unit_head, (unsigned long long)VKS_CHUNK);
        const uint32_t e = (S->fetched / VKS_CHUNK) % VKS_RING;
        const unsigned long long b = u / n_pixels;
        const uint32_t p = (uint32_t)(u - b * n_pixels);
        S->chunk_b[e] = b > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)b;
        S->chunk_y[e] = p / width;
        S->chunk_x[e] = p - (p / width) * width;
        S->fetched += VKS_CHUNK;
    }
}
// Start the sample `s` of pixel (x, y) in `slot`: Camera::get_ray, depth 1 (src/main.rs:187-190).
VKD void staged_begin_sample_xy(const StagedCtx& C, uint32_t slot, uint32_t x, uint32_t y, uint32_t s) {
    const uint32_t pixel = y * C.a.width + x; // i = y*width + x, row 0 = bottom (src/main.rs:182-183)
    PathRng rng;
    rng.pixel = pixel;
    rng.sample = s;
    rng.key = make_uint2(C.a.seed_lo, C.a.seed_hi);
    float3 o, d;
    float time;
    camera_get_ray(C.cam, rng, x, y, C.a.width, C.a.height, o, d, time);
    StagedShared& S = C.S;
    S.ro[slot] = make_float4(o.x, o.y, o.z, time);
    S.rd[slot] = make_float4(d.x, d.y, d.z, __uint_as_float(1u)); // ray_color(ray, .., 1)
    S.bt[slot] = make_float4(1.0f, 1.0f, 1.0f, __uint_as_float(s));
    S.hp[slot].w = pixel;
}
VKD void staged_begin_sample(const StagedCtx& C, uint32_t slot, uint32_t pixel, uint32_t s) { // (multi-sample units only)
    staged_begin_sample_xy(C, slot, pixel % C.a.width, pixel / C.a.width, s);
}
// Warp-synchronous: the lanes with want == true take the CTA's next units (one shared atomic per
// warp) and start their first sample; a lane whose unit is past the end leaves its slot idle.
VKD void staged_take_units(const StagedCtx& C, bool want, uint32_t slot, uint32_t lane) {
    const uint32_t m = __ballot_sync(0xFFFFFFFFu, want);
    if (m == 0u) return;
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(&C.S.next_unit, (uint32_t)__popc(m));
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    if (!want) return;
    const uint32_t n = base + __popc(m & ((1u << lane) - 1u));
    // unit = chunk start + n % CHUNK, walked in (sample block, row, column) without dividing
    const uint32_t e = (n / VKS_CHUNK) % VKS_RING;
    uint32_t b = C.S.chunk_b[e], y = C.S.chunk_y[e], x = C.S.chunk_x[e] + (n % VKS_CHUNK);
    while (x >= C.a.width) {
        x -= C.a.width;
        ++y;
    }
    while (y >= C.a.height) {
        y -= C.a.height;
        ++b;
    }
    if (b >= C.a.spp_count) { // past the last unit: the queue is empty
        C.S.rd[slot].w = __uint_as_float(0u);
        C.S.no_units = 1u;
        return;
    }
    staged_begin_sample_xy(C, slot, x, y, C.a.spp_begin + b);
}

// The sort: one __match_any_sync groups the lanes of the warp by shading class; the first lane of each
// group reserves the group's entries in the class list with one shared-memory atomic (full warp, converged).
VKD void staged_sort_append(StagedShared& S, uint32_t* cnt, uint32_t cls, uint32_t slot, uint32_t lane, uint32_t lanes_below) {
    const uint32_t m = __match_any_sync(0xFFFFFFFFu, cls); // lanes of my class (idle lanes form their own group)
    const uint32_t leader = __ffs(m) - 1u;
    uint32_t base = 0;
    if (lane == leader && cls != VKS_C_IDLE) base = atomicAdd(&cnt[cls], (uint32_t)__popc(m));
    base = __shfl_sync(0xFFFFFFFFu, base, leader);
    if (cls != VKS_C_IDLE) S.list[cls][base + __popc(m & lanes_below)] = (uint16_t)slot;
}

template <bool FLAT, bool MEDIA>
VKD void staged_body(const DScene& sc, const FlatProgram* flat, const DCamera& cam, const RenderArgs& a, const RenderBuffers& buf,
                     unsigned long long* unit_head) {
    extern __shared__ __align__(16) unsigned char vks_raw[];
    StagedShared& S = *reinterpret_cast<StagedShared*>(vks_raw);
    const uint32_t tid = threadIdx.x, lane = tid & 31u, lanes_below = (1u << lane) - 1u;
    const StagedCtx C = {cam, a, S, a.width * a.height, (unsigned long long)(a.width * a.height) * a.spp_count, unit_head};
    uint32_t n_rays = 0, n_drop = 0, n_nodes = 0, n_prims = 0;
    uint32_t dbg_iters = 0, dbg_sparse = 0; // iterations run; iterations with fewer than N/8 live slots

    if (tid == 0) {
        S.n_late = 0u;
        S.no_units = 0u;
        S.next_unit = 0u;
        S.fetched = 0u;
        staged_refill(&S, unit_head, a.width, C.n_pixels);
#ifdef VKS_DEBUG_TIMES
        unsigned long long t0;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        atomicMin(&buf.counters[5], t0); // first CTA start / first and last CTA end (ns), scripts/cta_spread.py
#endif
    }
    if (tid < 2 * VKS_CLASSES) (&S.cnt[0][0])[tid] = 0u;
    __syncthreads();
#pragma unroll 1
    for (uint32_t r = 0; r < VKS_ROUNDS; ++r) staged_take_units(C, true, r * VKS_T + tid, lane); // generate
    __syncthreads();
#pragma unroll 1
    for (uint32_t iter = 0;; ++iter) {
        uint32_t* cnt = S.cnt[iter & 1u];
        if (tid == 0) staged_refill(&S, unit_head, a.width, C.n_pixels); // consumed by the shade stage, after the barrier
        // ---- extend + classify (slot == thread + round * T: conflict-free shared-memory access) -------
        if (FLAT) {
#pragma unroll 1
            for (uint32_t r0 = 0; r0 < VKS_ROUNDS; r0 += VKS_K) {
                float3 o[VKS_K], d[VKS_K];
                float tm[VKS_K], best_t[VKS_K];
                bool live[VKS_K];
                uint32_t best_hit[VKS_K];
                MediumXi xi[VKS_K];
#pragma unroll
                for (int q = 0; q < VKS_K; ++q) {
                    const uint32_t slot = (r0 + q) * VKS_T + tid;
                    const float4 ro = S.ro[slot], rd = S.rd[slot];
                    o[q] = f3(ro);
                    d[q] = f3(rd);
                    tm[q] = ro.w;
                    live[q] = __float_as_uint(rd.w) != 0u;
                    xi[q].table = nullptr;
                    xi[q].depth = __float_as_uint(rd.w);
                    xi[q].rng.key = make_uint2(a.seed_lo, a.seed_hi);
                    xi[q].rng.pixel = MEDIA ? S.hp[slot].w : 0u;
                    xi[q].rng.sample = MEDIA ? __float_as_uint(S.bt[slot].w) : 0u;
                }
                bool any_live = false;
#pragma unroll
                for (int q = 0; q < VKS_K; ++q) any_live = any_live || live[q];
                if (!__any_sync(0xFFFFFFFFu, any_live)) continue; // draining pool: nothing for this warp in these rounds
                trace_flat_k<VKS_K, MEDIA>(sc, *flat, o, d, tm, live, 0.001f, xi, best_t, best_hit); // src/main.rs:130
#pragma unroll
                for (int q = 0; q < VKS_K; ++q) {
                    const uint32_t slot = (r0 + q) * VKS_T + tid;
                    uint32_t cls = VKS_C_IDLE;
                    if (live[q]) {
                        ++n_rays;
                        n_prims += flat->n;
                        uint32_t prim = VK_REF_NONE, hi = 0u;
                        cls = VKS_C_TERMINATE;
                        if (best_hit[q] != 0xFFFFFFFFu) {
                            const FlatHit& fh = flat->hits[VKF_HIT_INDEX(best_hit[q])];
                            prim = fh.prim & ~VKD_DUP;
                            const uint32_t inst = fh.inst;
                            hi = (inst ? (0x80000000u | VKD_INDEX(inst)) : 0u) | (fh.face << 28);
                            cls = fh.cls;
                        }
                        S.hp[slot].x = __float_as_uint(best_t[q]);
                        S.hp[slot].y = prim;
                        S.hp[slot].z = hi;
                    }
                    staged_sort_append(S, cnt, cls, slot, lane, lanes_below);
                }
            }
        } else {
#pragma unroll 1
            for (uint32_t r = 0; r < VKS_ROUNDS; ++r) {
                const uint32_t slot = r * VKS_T + tid;
                const float4 rd = S.rd[slot];
                const uint32_t depth = __float_as_uint(rd.w);
                uint32_t cls = VKS_C_IDLE;
                if (depth) {
                    const float4 ro = S.ro[slot];
                    MediumXi xi;
                    xi.table = nullptr;
                    xi.depth = depth;
                    xi.rng.key = make_uint2(a.seed_lo, a.seed_hi);
                    xi.rng.pixel = MEDIA ? S.hp[slot].w : 0u;
                    xi.rng.sample = MEDIA ? __float_as_uint(S.bt[slot].w) : 0u;
                    TraceCounters tc = {0u, 0u};
                    const TraceHit h = trace<MEDIA>(sc, f3(ro), f3(rd), ro.w, 0.001f, CUDART_INF_F, xi, tc); // src/main.rs:130
                    ++n_rays;
                    n_nodes += tc.nodes;
                    n_prims += tc.prims;
                    S.hp[slot].x = __float_as_uint(h.t);
                    S.hp[slot].y = h.prim;
                    S.hp[slot].z = (h.inst ? (0x80000000u | VKD_INDEX(h.inst)) : 0u) | (h.face << 28);
                    cls = h.prim == VK_REF_NONE ? (uint32_t)VKS_C_TERMINATE : staged_class_of(sc, h.prim);
                }
                staged_sort_append(S, cnt, cls, slot, lane, lanes_below);
            }
        }
        if (tid < VKS_CLASSES) S.cnt[(iter & 1u) ^ 1u][tid] = 0u; // next iteration's counters (last read before the previous barrier)
        __syncthreads();
        const uint32_t c0 = cnt[0], c1 = cnt[1], c2 = cnt[2], c3 = cnt[3];
        const uint32_t o1 = c0, o2 = c0 + c1, o3 = c0 + c1 + c2, n_live = o3 + c3;
        if (n_live == 0u) break; // pool drained and no unit left (uniform: every thread reads the same counters)
        ++dbg_iters;
        if (n_live < VKS_N / 8) ++dbg_sparse;
        // One slot of the shade stage: resolve + scatter (src/main.rs:131-149); a finished sample goes through the
        // NaN/Inf filter into its plane (:191-194).  Reports whether the sample ended and what the slot needs next.
        auto shade_slot = [&](uint32_t slot, bool& ended, bool& new_unit, uint32_t& next_sample) {
                const float4 ro = S.ro[slot], rd = S.rd[slot], bt = S.bt[slot];
                const uint4 hp = S.hp[slot];
                float3 o = f3(ro), d = f3(rd), beta = f3(bt), L = f3(0.0f, 0.0f, 0.0f);
                float time = ro.w;
                uint32_t depth = __float_as_uint(rd.w);
                const uint32_t pixel = hp.w, sample = __float_as_uint(bt.w);
                const uint32_t prim = hp.y;
                bool alive, valid = true;
                if (prim == VK_REF_NONE) {
                    L = beta * miss_color(a, d); // src/main.rs:151
                    alive = false;
                } else {
                    PathRng rng;
                    rng.pixel = pixel;
                    rng.sample = sample;
                    rng.key = make_uint2(a.seed_lo, a.seed_hi);
                    const uint32_t hi = hp.z;
                    TraceHit h;
                    h.t = __uint_as_float(hp.x);
                    h.prim = prim;
                    h.inst = (hi & 0x80000000u) ? (((uint32_t)VK_T_XFORM << 28) | (hi & 0x07FFFFFFu)) : 0u;
                    h.face = (hi >> 28) & 7u;
                    HitRecD rec;
                    resolve_hit(sc, h, o, d, time, false, rec);
                    alive = shade(sc, rec, rng, depth, o, d, time, beta, L, valid);
                    if (alive && ++depth > a.max_depth) alive = false; // `depth > MAX_DEPTH` -> 0 (src/main.rs:126)
                    if (alive && !(finite3(d) && finite3(o))) {          // the reference's sample is NaN here (see vk_kernels.cu)
                        valid = false;
                        alive = false;
                    }
                }
                if (alive) {
                    S.ro[slot] = make_float4(o.x, o.y, o.z, time);
                    S.rd[slot] = make_float4(d.x, d.y, d.z, __uint_as_float(depth));
                    S.bt[slot] = make_float4(beta.x, beta.y, beta.z, bt.w);
                } else { // sample finished: NaN/Inf filter of src/main.rs:191-194, into the pixel's integer accumulators
                    const bool keep = valid && finite3(L);
                    if (!keep) ++n_drop;
                    else accumulate_sample(buf, pixel, L);
                    ended = true;
                    next_sample = sample + 1u;
                    new_unit = true; // a unit is one sample
                    S.rd[slot].w = __uint_as_float(0u); // idle until regenerated
                }
        };
        auto slot_of = [&](uint32_t j) -> uint32_t {
            return j < o1 ? S.list[0][j] : (j < o2 ? S.list[1][j - o1] : (j < o3 ? S.list[2][j - o2] : S.list[3][j - o3]));
        };
        // ---- drain: the queue is empty and one warp's worth of paths is left (the long ones: up to max_depth
        // segments each).  A staged iteration costs ~5 us however few slots are live, so warp 0 finishes them
        // alone, one lane per path, trace and shade back to back without barriers; the other warps leave.
        if (S.no_units != 0u && n_live <= 32u) {
            if (tid < 32u && lane < n_live) {
                const uint32_t slot = slot_of(lane);
#pragma unroll 1
                for (;;) {
                    bool ended = false, new_unit = false;
                    uint32_t next_sample = 0;
                    shade_slot(slot, ended, new_unit, next_sample);
                    if (ended) {
                        if (new_unit) break; // no unit left: the slot is done
                        staged_begin_sample(C, slot, S.hp[slot].w, next_sample);
                    }
                    const float4 ro = S.ro[slot], rd = S.rd[slot];
                    MediumXi xi;
                    xi.table = nullptr;
                    xi.depth = __float_as_uint(rd.w);
                    xi.rng.key = make_uint2(a.seed_lo, a.seed_hi);
                    xi.rng.pixel = S.hp[slot].w;
                    xi.rng.sample = __float_as_uint(S.bt[slot].w);
                    TraceCounters tc = {0u, 0u};
                    const TraceHit h = FLAT ? trace_flat<MEDIA>(sc, *flat, f3(ro), f3(rd), ro.w, 0.001f, CUDART_INF_F, xi, tc)
                                            : trace<MEDIA>(sc, f3(ro), f3(rd), ro.w, 0.001f, CUDART_INF_F, xi, tc); // src/main.rs:130
                    ++n_rays;
                    n_nodes += tc.nodes;
                    n_prims += tc.prims;
                    S.hp[slot].x = __float_as_uint(h.t);
                    S.hp[slot].y = h.prim;
                    S.hp[slot].z = (h.inst ? (0x80000000u | VKD_INDEX(h.inst)) : 0u) | (h.face << 28);
                }
            }
            break;
        }
        // ---- shade: a warp's 32 consecutive entries are one class (except at the 3 class borders) ------
#pragma unroll 1
        for (uint32_t j0 = 0; j0 < n_live; j0 += VKS_T) {
            const uint32_t j = j0 + tid;
            bool new_unit = false, ended = false;
            uint32_t slot = 0, next_sample = 0;
            if (j < n_live) {
                slot = slot_of(j);
                shade_slot(slot, ended, new_unit, next_sample);
            }
            // Regenerate in place (src/main.rs:187-190) when at least half the warp ended -- the emitter /
            // miss class does, every lane -- otherwise queue the slot: in the other classes only a few
            // lanes end (a light-sampled direction below the surface has weight 0), and the 300
            // instructions of Philox + camera ray would run with those few lanes active.
            const uint32_t m_end = __ballot_sync(0xFFFFFFFFu, ended);
            if ((uint32_t)__popc(m_end) >= 16u) {
                staged_take_units(C, ended && new_unit, slot, lane);
                if (ended && !new_unit) staged_begin_sample(C, slot, S.hp[slot].w, next_sample);
            } else if (m_end) {
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(&S.n_late, (uint32_t)__popc(m_end));
                base = __shfl_sync(0xFFFFFFFFu, base, 0);
                if (ended) S.late[base + __popc(m_end & lanes_below)] = (uint16_t)(slot | (new_unit ? 0x8000u : 0u));
            }
        }
        __syncthreads();
        {   // the queued regenerations, with full warps
            const uint32_t n_late = S.n_late;
#pragma unroll 1
            for (uint32_t k0 = 0; k0 < n_late; k0 += VKS_T) {
                const uint32_t k = k0 + tid;
                const bool have = k < n_late;
                const uint32_t e = have ? S.late[k] : 0u, slot = e & 0x7FFFu;
                const bool new_unit = have && (e & 0x8000u);
                staged_take_units(C, new_unit, slot, lane);
                if (have && !new_unit) staged_begin_sample(C, slot, S.hp[slot].w, __float_as_uint(S.bt[slot].w) + 1u);
            }
        }
        __syncthreads();
        if (tid == 0) S.n_late = 0u; // next written in the shade stage, two barriers away
    }
#ifdef VKS_DEBUG_TIMES
    __shared__ unsigned long long s_cta_rays;
    if (tid == 0) s_cta_rays = 0ull;
    __syncthreads();
#endif
    unsigned long long w_rays = n_rays, w_drop = n_drop, w_nodes = n_nodes, w_prims = n_prims;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        w_rays += __shfl_xor_sync(0xFFFFFFFFu, w_rays, off);
        w_drop += __shfl_xor_sync(0xFFFFFFFFu, w_drop, off);
        w_nodes += __shfl_xor_sync(0xFFFFFFFFu, w_nodes, off);
        w_prims += __shfl_xor_sync(0xFFFFFFFFu, w_prims, off);
    }
    if (lane == 0) {
        atomicAdd(&buf.counters[3], w_nodes);
        atomicAdd(&buf.counters[4], w_prims);
        atomicAdd(&buf.counters[0], w_rays);
        if (w_drop) atomicAdd(&buf.counters[1], w_drop);
#ifdef VKS_DEBUG_TIMES
        atomicAdd(&s_cta_rays, w_rays);
#endif
    }
#ifdef VKS_DEBUG_TIMES
    __syncthreads();
    if (tid == 0) {
        unsigned long long t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        atomicMin(&buf.counters[6], t1);
        atomicMax(&buf.counters[7], t1);
        if (buf.debug && blockIdx.x < VK_DEBUG_CTAS) {
            uint32_t smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            buf.debug[4 * blockIdx.x + 0] = t1;
            buf.debug[4 * blockIdx.x + 1] = s_cta_rays;
            buf.debug[4 * blockIdx.x + 2] = smid;
            buf.debug[4 * blockIdx.x + 3] = dbg_iters | ((unsigned long long)dbg_sparse << 32);
        }
    }
#else
    (void)dbg_iters, (void)dbg_sparse;
#endif
}

template <bool MEDIA>
__global__ void __launch_bounds__(VKS_T, VKS_MINB) k_staged(const DScene sc, const DCamera cam, const RenderArgs a, const RenderBuffers buf,
                                                     unsigned long long* unit_head) {
    staged_body<false, MEDIA>(sc, nullptr, cam, a, buf, unit_head);
}
template <bool MEDIA>
__global__ void __launch_bounds__(VKS_T, VKS_MINB) k_staged_flat(const DScene sc, const __grid_constant__ FlatProgram flat, const DCamera cam,
                                                          const RenderArgs a, const RenderBuffers buf, unsigned long long* unit_head) {
    staged_body<true, MEDIA>(sc, &flat, cam, a, buf, unit_head);
}

template <class K> static cudaError_t staged_prepare(K kernel, int* blocks_per_sm) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(StagedShared));
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, kernel, VKS_T, sizeof(StagedShared));
}

// grid = sm_count * resident CTAs; *unit_head must be zero on the stream before the launch
cudaError_t launch_staged(const DScene& sc, const FlatProgram* flat, const DCamera& cam, const RenderArgs& a, const RenderBuffers& b,
                          unsigned long long* unit_head, int sm_count, cudaStream_t st) {
    int bps = 0;
    cudaError_t e;
    const size_t smem = sizeof(StagedShared);
    if (flat && flat->n) {
        if (sc.has_media) {
            if ((e = staged_prepare(k_staged_flat<true>, &bps)) != cudaSuccess) return e;
            k_staged_flat<true><<<sm_count * (bps < 1 ? 1 : bps), VKS_T, smem, st>>>(sc, *flat, cam, a, b, unit_head);
        } else {
            if ((e = staged_prepare(k_staged_flat<false>, &bps)) != cudaSuccess) return e;
            k_staged_flat<false><<<sm_count * (bps < 1 ? 1 : bps), VKS_T, smem, st>>>(sc, *flat, cam, a, b, unit_head);
        }
    } else {
        if (sc.has_media) {
            if ((e = staged_prepare(k_staged<true>, &bps)) != cudaSuccess) return e;
            k_staged<true><<<sm_count * (bps < 1 ? 1 : bps), VKS_T, smem, st>>>(sc, cam, a, b, unit_head);
        } else {
            if ((e = staged_prepare(k_staged<false>, &bps)) != cudaSuccess) return e;
            k_staged<false><<<sm_count * (bps < 1 ? 1 : bps), VKS_T, smem, st>>>(sc, cam, a, b, unit_head);
        }
    }
    return cudaGetLastError();
}

} // namespace VK_NS
