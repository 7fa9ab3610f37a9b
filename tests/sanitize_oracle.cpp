// sanitize_oracle.cpp -- the CHECKER under AddressSanitizer + UndefinedBehaviorSanitizer (+ float-cast-overflow): oracle/oracle.cpp
// is what every parity test trusts, so an out-of-bounds read, an uninitialised-looking branch or a float -> integer cast
// whose C++ meaning differs from Rust's saturating `as` (NaN -> 0, too large -> MAX) would silently bend the reference
// side of the comparisons.  Built and run by tests/test_sanitizers.py; no GPU, no product kernel.
//
// Per scene: rebuild the object graph from the lowered description, harvest the ray segments of a few samples, intersect
// them (also with injected medium variates), render a small frame with HEAD's integrator (the legacy one for the scenes
// without a light list), run the known-answer hooks' scene-free entry, and push awkward records through orc_eval_batch.
// usage: sanitize_oracle [assets_dir]
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../oracle/oracle.h"
#include "vecchio_host.h"

static int fail(const char* what, const char* why) {
    std::fprintf(stderr, "sanitize_oracle: %s: %s\n", what, why);
    return 1;
}

int main(int argc, char** argv) {
    const char* assets = argc > 1 ? argv[1] : "assets";
    struct Job { const char* name; uint32_t param; uint32_t flags; uint32_t depth; };
    const uint32_t LEGACY = VK_FLAG_LEGACY_SCATTER | VK_FLAG_SKY_BACKGROUND;
    const Job jobs[] = {{"cornell_box", 0, 0, 100},   {"cornell_smoke", 0, 0, 100},       {"final_scene", 0, 0, 100},  {"random_spheres_demo", 0, 0, 50},
                        {"bowser_demo", 0, 0, 50},    {"perlin_demo", 0, 0, 50},          {"balls_demo", 0, 0, 50},    {"api_surface_demo", 0, 0, 50},
                        {"random_spheres_cover", 0, LEGACY, 50}, {"book1_cover", 0, LEGACY, 50}, {"stress_spheres", 24, 0, 50},
                        {"furnace_demo", 0, 0, 100},  {"furnace_demo", 1, 0, 100},        {"furnace_demo", 2, 0, 100}, {"furnace_demo", 3, 0, 100},
                        {"cornell_box", 0, VK_FLAG_LEGACY_SCATTER, 50}};
    unsigned long long rays_total = 0, paths_total = 0;
    for (const Job& job : jobs) {
        vkh_scene* sc = nullptr;
        if (vkh_scene_build(job.name, 1, assets, job.param, &sc) != VK_OK) return fail(job.name, vkh_last_error());
        const vk_scene_desc* d = vkh_scene_desc(sc);
        vk_camera cam;
        if (!vkh_scene_next_camera(sc, &cam)) return fail(job.name, "no camera");
        orc_scene* o = nullptr;
        if (orc_scene_create(d, &o) != VK_OK) return fail(job.name, orc_last_error());

        const uint32_t W = 40, H = (uint32_t)((float)W / vkh_scene_aspect_ratio(sc));
        std::vector<vk_ray> rays(6000);
        if (d->n_lights) { // ray_color needs the light list (HEAD); the cover scenes are legacy-only
            rays.resize(orc_harvest_rays(o, &cam, W, H, job.depth, 7, rays.size(), rays.data()));
        } else {
            rays.resize(64);
            for (size_t i = 0; i < rays.size(); ++i) {
                vk_ray r{};
                for (int k = 0; k < 3; ++k) r.origin[k] = cam.origin[k];
                for (int k = 0; k < 3; ++k)
                    r.direction[k] = cam.lower_left_corner[k] + (float)(i % 8) / 7.0f * cam.horizontal[k] + (float)(i / 8) / 7.0f * cam.vertical[k] - cam.origin[k];
                r.time = 0.5f; r.tmin = 0.001f; r.tmax = INFINITY;
                rays[i] = r;
            }
        }
        // degenerate rays as well: zero direction, NaN, infinities
        const float bad[] = {0.0f, NAN, INFINITY, -INFINITY, 1e-38f};
        for (int k = 0; k < 10; ++k) {
            vk_ray r = rays[k % rays.size()];
            r.direction[k % 3] = bad[k % 5];
            if (k >= 5) r.origin[(k + 1) % 3] = bad[(k + 2) % 5];
            rays.push_back(r);
        }
        std::vector<vk_hit> hits(rays.size());
        if (orc_intersect(o, rays.data(), rays.size(), nullptr, hits.data()) != VK_OK) return fail(job.name, orc_last_error());
        if (d->n_media && d->n_media <= VK_MEDIUM_XI_SLOTS / 2) {
            std::vector<float> xi(rays.size() * VK_MEDIUM_XI_SLOTS);
            for (size_t i = 0; i < xi.size(); ++i) xi[i] = (float)((i * 2654435761u) % 1000u + 1u) / 1001.0f;
            if (orc_intersect(o, rays.data(), rays.size(), xi.data(), hits.data()) != VK_OK) return fail(job.name, orc_last_error());
        }
        rays_total += rays.size();

        vk_render_params p{};
        p.width = W; p.height = H; p.spp = 4; p.max_depth = job.depth; p.seed = 3; p.flags = job.flags;
        std::vector<float> rgb((size_t)W * H * 3), sq(rgb.size());
        orc_stats st{};
        const int rc = orc_render(o, &cam, &p, rgb.data(), sq.data(), &st, 2);
        if (rc != VK_OK && !(d->n_lights == 0 && !(job.flags & VK_FLAG_LEGACY_SCATTER))) return fail(job.name, orc_last_error());
        paths_total += st.paths;
        // a slice of the samples, and a frame with a lens
        p.spp = 8; p.spp_begin = 3; p.spp_count = 2;
        if (orc_render(o, &cam, &p, rgb.data(), nullptr, &st, 1) != VK_OK) return fail(job.name, orc_last_error());
        vk_camera lens = cam;
        lens.lens_radius = 0.3f;
        p.spp_begin = 0; p.spp_count = 0; p.spp = 2;
        if (orc_render(o, &lens, &p, rgb.data(), nullptr, &st, 2) != VK_OK) return fail(job.name, orc_last_error());

        // the shading-side hook: every material / texture / light of the scene with ordinary and awkward inputs
        std::vector<vk_eval> recs;
        const float vals[] = {0.0f, 1.0f, -1.0f, 0.5f, NAN, INFINITY, 1e30f, -1e30f, 1e-30f};
        for (uint32_t op = VK_EVAL_BOUNCE; op <= VK_EVAL_LIGHT_RANDOM; ++op) {
            const uint32_t n_index = op == VK_EVAL_TEXTURE ? d->n_textures : op <= VK_EVAL_BOUNCE_LEGACY ? d->n_materials : op == VK_EVAL_LIGHT_RANDOM ? d->n_lights : 1;
            if ((op == VK_EVAL_LIGHTS_PDF || op == VK_EVAL_LIGHT_RANDOM || op == VK_EVAL_BOUNCE) && d->n_lights == 0) continue;
            for (uint32_t ix = 0; ix < n_index && ix < 64; ++ix)
                for (int v = 0; v < 12; ++v) {
                    vk_eval e{};
                    e.op = op; e.index = ix;
                    const float a = v < 3 ? 0.3f + 0.2f * v : vals[(v + ix) % 9];
                    e.ray_o[0] = 1; e.ray_o[1] = 2; e.ray_o[2] = 3; e.ray_d[0] = 0.2f; e.ray_d[1] = -0.9f; e.ray_d[2] = a; e.ray_time = 0.25f;
                    e.p[0] = 10 * a; e.p[1] = 5; e.p[2] = -3; e.normal[0] = 0; e.normal[1] = 1; e.normal[2] = v < 3 ? 0.0f : vals[(v * 5 + 1) % 9];
                    e.t = 1.5f; e.u = v < 3 ? 0.25f * v : vals[(v + 3) % 9]; e.v = v < 3 ? 0.7f : vals[(v + 5) % 9]; e.front = v & 1;
                    e.dir[0] = 0.1f; e.dir[1] = 1.0f; e.dir[2] = v < 3 ? 0.2f : vals[(v + 7) % 9];
                    for (int k = 0; k < 5; ++k) e.xi[k] = v == 3 ? 0u : v == 4 ? 0xFFFFFFFFu : (uint32_t)(2654435761u * (uint32_t)(v * 5 + k + ix + 1));
                    recs.push_back(e);
                }
        }
        if (!recs.empty() && orc_eval_batch(o, recs.data(), recs.size()) != VK_OK) return fail(job.name, orc_last_error());
        orc_scene_free(o);
        vkh_scene_free(sc);
    }
    std::printf("sanitize_oracle ok: %llu rays intersected, %llu paths rendered\n", rays_total, paths_total);
    return 0;
}
