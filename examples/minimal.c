/* minimal.c -- the C ABI from plain C: build the Cornell box with the host front end, upload it, render one
 * small frame, write the reference's P3 file.  What a binding in any language does, in the order it does it.
 *
 *   gcc -std=c11 -Iinclude examples/minimal.c -Lvecchio_b200/lib -lvecchio_host -lvecchio_gpu \
 *       -Wl,-rpath,$PWD/vecchio_b200/lib -o /tmp/minimal && /tmp/minimal out.ppm
 *
 * Exit code 0 = frame written; 1 = a library call failed (message on stderr; without a CUDA device that is
 * vk_create: there is no CPU path). */
#include <stdio.h>
#include <stdlib.h>

#include "vecchio_gpu.h"
#include "vecchio_host.h"

int main(int argc, char** argv) {
    const char* out_path = argc > 1 ? argv[1] : "output_0000.ppm";
    const uint32_t width = 200, spp = 64, max_depth = 100;

    vkh_scene* scene = NULL;
    if (vkh_scene_build("cornell_box", 1, NULL, 0, &scene) != VK_OK) { /* cornell_box() + BVHNode::new, src/main.rs:159-168 */
        fprintf(stderr, "scene: %s\n", vkh_last_error());
        return 1;
    }
    const uint32_t height = (uint32_t)((float)width / vkh_scene_aspect_ratio(scene)); /* src/main.rs:172 */

    vk_scene_info info;
    char why[256];
    if (vk_scene_check(vkh_scene_desc(scene), &info, why, sizeof why) != VK_OK) { /* host only: no device needed */
        fprintf(stderr, "scene check: %s\n", why);
        vkh_scene_free(scene);
        return 1;
    }
    fprintf(stderr, "layout: %u flat entries in %u segments, %u wide nodes, stack %u\n", info.flat_entries,
            info.flat_segments, info.wide_nodes, info.stack_need);

    vk_ctx* ctx = NULL;
    if (vk_create(0, &ctx) != VK_OK) {
        fprintf(stderr, "vk_create: %s\n", vk_last_error(NULL));
        vkh_scene_free(scene);
        return 1;
    }
    int rc = vk_scene_upload(ctx, vkh_scene_desc(scene));
    vk_camera cam;
    if (rc == VK_OK && !vkh_scene_next_camera(scene, &cam)) rc = VK_ERR_INVALID; /* config.cam_iter.next(), src/main.rs:176 */

    uint8_t* rgb8 = (uint8_t*)malloc((size_t)width * height * 3);
    vk_stats st;
    if (rc == VK_OK && rgb8) {
        vk_render_params p = {0};
        p.width = width;
        p.height = height;
        p.spp = spp;
        p.max_depth = max_depth;
        p.seed = 1;
        rc = vk_render_rgb8(ctx, &cam, &p, rgb8, &st); /* src/main.rs:181-198 and the conversion of :201-214 */
    }
    if (rc != VK_OK) {
        fprintf(stderr, "render: %s\n", vk_last_error(ctx));
    } else if (vkh_write_ppm(out_path, rgb8, width, height) != VK_OK) {
        fprintf(stderr, "write: %s\n", vkh_last_error());
        rc = VK_ERR_INVALID;
    } else {
        fprintf(stderr, "Wrote frame %s: %llu paths, %llu rays, %.2f ms of kernels\n", out_path,
                (unsigned long long)st.paths, (unsigned long long)st.rays, st.ms_kernels);
    }
    free(rgb8);
    vk_destroy(ctx);
    vkh_scene_free(scene);
    return rc == VK_OK ? 0 : 1;
}
