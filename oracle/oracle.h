/*
 * oracle.h -- C entry points of the CPU oracle.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library, and only as the checker or the
 * reported CPU baseline.  The product path (libvecchio_gpu.so) never links or calls it.
 *
 * The oracle is a C++ restatement of the reference's Rust (browserdotsys/vecchio), function by
 * function, in oracle.cpp.  PARITY PINNING: the reference has no tests, golden vectors or
 * fixtures; rustc is absent so the reference cannot run here (no oracle/_ref).  The oracle is
 * pinned by (i) the analytic known-answer vectors derived from the cited formulas, (ii) second,
 * independent numpy restatements of its trickier functions (tests/test_oracle.py), (iii) closed
 * forms of ray_color in a furnace scene, and (iv) region means of the three renders the reference
 * publishes, sample/{therestofyourlife,thenextweek,inoneweekend}.png (tests/golden/).  By the
 * task's rule this is "parity unpinned" by reference tests; it is pinned only by those images
 * and the source.
 */
#ifndef VECCHIO_ORACLE_H
#define VECCHIO_ORACLE_H

#include "../include/vecchio_gpu.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_scene orc_scene;

typedef struct orc_stats {
    uint64_t paths, rays, dropped_samples;
    /* traversal / shading event counts on the reference's own traversal order */
    uint64_t n_node, n_sph_rej, n_sph_acc, n_msph, n_rect_rej, n_rect_acc, n_box, n_translate, n_rotate,
        n_medium, n_texel, n_perlin, n_diffuse, n_dielectric, n_metal, n_emit_or_miss, n_light_pdf;
    uint64_t rays_live; /* rays traced while the path weight was still non-zero and finite */
    double seconds;
    int threads;
} orc_stats;

/* Rebuild the reference's object graph (Arc<dyn Hittable> etc.) from a flattened scene. */
int orc_scene_create(const vk_scene_desc* desc, orc_scene** out);
void orc_scene_free(orc_scene* s);
const char* orc_last_error(void);

/* world.hit(&r, tmin, tmax) for each ray (src/accel.rs:58-83 and below). */
int orc_intersect(const orc_scene* s, const vk_ray* rays, size_t n, const float* medium_xi, vk_hit* out);

/* The sample loop src/main.rs:181-198 (OpenMP over pixels = rayon par_iter_mut).
 * out_rgb: mean over params->spp; out_sumsq: per-channel sum of squares of kept samples. */
int orc_render(const orc_scene* s, const vk_camera* cam, const vk_render_params* params, float* out_rgb,
               float* out_sumsq, orc_stats* stats, int n_threads);

/* Every ray segment ray_color() traces for a few samples per pixel (camera + bounce rays, with
 * their unnormalised directions) -- the secondary-ray half of the hit-parity batches. */
size_t orc_harvest_rays(const orc_scene* s, const vk_camera* cam, uint32_t width, uint32_t height,
                        uint32_t max_depth, uint64_t seed, size_t max_rays, vk_ray* out);

/* The oracle's side of vk_eval_batch: the same records through the restated reference code with the same
 * variates (see the comment at its definition for how the words are put in the reference's draw order). */
int orc_eval_batch(const orc_scene* s, vk_eval* recs, size_t n);

/* Known-answer hooks for the unit KATs; returns number of outputs written or -1. */
int orc_kat(const orc_scene* s, const char* name, const float* in, int n_in, float* out, int n_out);

#ifdef __cplusplus
}
#endif
#endif
