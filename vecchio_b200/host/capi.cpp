// capi.cpp -- extern "C" view of the host front end (include/vecchio_host.h).
#include "../../include/vecchio_host.h"
#include "vecchio.hpp"

using namespace vecchio;

struct vkh_scene {
    std::unique_ptr<LoweredScene> scene;
    vk_scene_desc desc;
};

static thread_local std::string g_err;

extern "C" {

int vkh_scene_build(const char* name, uint64_t seed, const char* assets_dir, uint32_t param, vkh_scene** out) {
    if (!name || !out) {
        g_err = "vkh_scene_build: null argument";
        return VK_ERR_INVALID;
    }
    try {
        g_assets_dir = assets_dir ? assets_dir : "assets";
        seed_thread_rng(seed);
        std::string n(name);
        SceneConfig cfg;
        if (n == "balls_demo") cfg = balls_demo();
        else if (n == "random_spheres_demo") cfg = random_spheres_demo();
        else if (n == "random_spheres_cover") cfg = random_spheres_cover();
        else if (n == "api_surface_demo") cfg = api_surface_demo();
        else if (n == "perlin_demo") cfg = perlin_demo();
        else if (n == "bowser_demo") cfg = bowser_demo();
        else if (n == "cornell_box") cfg = cornell_box();
        else if (n == "final_scene") cfg = final_scene();
        else if (n == "cornell_smoke") cfg = cornell_smoke();
        else if (n == "stress_spheres") cfg = stress_spheres(param ? param : 1000);
        else {
            g_err = "Not a valid scene: " + n; // panic!("Not a valid scene") src/main.rs:166
            return VK_ERR_INVALID;
        }
        auto s = new vkh_scene;
        s->scene = lower_scene(std::move(cfg));
        s->desc = s->scene->low.desc();
        *out = s;
        return VK_OK;
    } catch (const LowerError& e) {
        g_err = e.what();
        return VK_ERR_UNSUPPORTED;
    } catch (const std::exception& e) {
        g_err = e.what();
        return VK_ERR_INVALID;
    }
}

void vkh_scene_free(vkh_scene* s) { delete s; }
const vk_scene_desc* vkh_scene_desc(const vkh_scene* s) { return s ? &s->desc : nullptr; }
float vkh_scene_aspect_ratio(const vkh_scene* s) { return s ? s->scene->aspect_ratio : 0.0f; }

int vkh_scene_next_camera(vkh_scene* s, vk_camera* out) {
    if (!s || !out || !s->scene->cam_iter) return 0;
    auto c = s->scene->cam_iter->next();
    if (!c) return 0;
    *out = c->lower();
    return 1;
}

void vkh_camera_new(const float lookfrom[3], const float lookat[3], const float vup[3], float vfov,
                    float aspect_ratio, float aperture, float focus_dist, float time0, float time1,
                    vk_camera* out) {
    Camera c = Camera::make(Vec3(lookfrom[0], lookfrom[1], lookfrom[2]), Vec3(lookat[0], lookat[1], lookat[2]),
                            Vec3(vup[0], vup[1], vup[2]), vfov, aspect_ratio, aperture, focus_dist, time0, time1);
    *out = c.lower();
}

long vkh_decode_png(const char* path, uint8_t* buf, size_t buf_len, uint32_t* width, uint32_t* height) {
    try {
        ImageTexture t(path);
        if (width) *width = (uint32_t)t.width;
        if (height) *height = (uint32_t)t.height;
        if (buf) {
            if (buf_len < t.buf.size()) {
                g_err = "vkh_decode_png: buffer too small";
                return -1;
            }
            std::memcpy(buf, t.buf.data(), t.buf.size());
        }
        return (long)t.buf.size();
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}

const char* vkh_last_error(void) { return g_err.c_str(); }

} // extern "C"
