/*
 * vecchio_host.h -- C view of the host front end (vecchio_b200/host/vecchio.hpp) for callers
 * that cannot include C++ (the Python tests and bench.py).  It builds a scene with the
 * reference's scene API (src/scene.rs builders), builds the world BVH like src/main.rs:168,
 * lowers both to a vk_scene_desc, and steps the scene's camera iterator (src/main.rs:176).
 * Nothing here computes a hit or a colour.
 */
#ifndef VECCHIO_HOST_H
#define VECCHIO_HOST_H

#include "vecchio_gpu.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vkh_scene vkh_scene;

/* name: balls_demo | random_spheres_demo | perlin_demo | bowser_demo | cornell_box |
 *       final_scene | cornell_smoke | stress_spheres (param = grid side, 1000 => 1M spheres) |
 *       api_surface_demo (SpecDiffuse, sphere and box lights) |
 *       random_spheres_cover (random_spheres_demo without its light: sky-lit, legacy integrator only)
 * seed: seeds the host RNG standing in for rand::thread_rng() (scene + BVH axis choices).
 * assets_dir: directory holding earthmap.png etc. (NULL => "assets"). */
int vkh_scene_build(const char* name, uint64_t seed, const char* assets_dir, uint32_t param, vkh_scene** out);
void vkh_scene_free(vkh_scene* s);
const vk_scene_desc* vkh_scene_desc(const vkh_scene* s);
float vkh_scene_aspect_ratio(const vkh_scene* s);
/* cam_iter.next(): 1 = camera written, 0 = iterator exhausted */
int vkh_scene_next_camera(vkh_scene* s, vk_camera* out);
/* Camera::new (src/main.rs:71-109) */
void vkh_camera_new(const float lookfrom[3], const float lookat[3], const float vup[3], float vfov,
                    float aspect_ratio, float aperture, float focus_dist, float time0, float time1,
                    vk_camera* out);
/* ImageTexture::new (src/material.rs:269-279): decoded RGB8 bytes; returns byte count or -1.
 * buf may be NULL to query the size. */
long vkh_decode_png(const char* path, uint8_t* buf, size_t buf_len, uint32_t* width, uint32_t* height);
const char* vkh_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
