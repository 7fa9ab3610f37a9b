// sanitize_host.cpp -- the HOST half of the product path under AddressSanitizer + UndefinedBehaviorSanitizer
// (built and run by tests/test_sanitizers.py; needs no GPU and makes no device call).
//
// What runs instrumented: the scene front end (vecchio_b200/host: the reference's scene builders, BVHNode::new, PNG decode,
// Perlin tables, every lower(), the camera iterators, the P3 writer) and everything vk_scene_upload does before it touches
// the device (vecchio_b200/csrc/vk_relayout.h behind vk_scene_check: the validator, the 4-wide BVH collapse, the flat-program
// builder, the shading records).  Two parts:
//   1. every scene the front end can build: build, lower, plan, walk the camera iterator, convert and write a frame, free;
//   2. corrupted descriptions: a deep copy of a lowered scene whose arrays are EXACTLY as long as their counts say (so that
//      an index the validator lets through lands in a redzone), one to three random corruptions per case -- references,
//      counts, record words, the root -- handed to vk_scene_check.  Any return code is fine; a sanitizer report is not.
// (compute-sanitizer for the device half is closed on this pool; the queue kernels check their own protocol in a debug
// build instead, tests/test_zz_warpq_selfcheck_gpu.py.)
// usage: sanitize_host [assets_dir [tmp.ppm [cases_per_scene [salt]]]]
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "vecchio_gpu.h"
#include "vecchio_host.h"

namespace {

struct Rng { // splitmix64
    uint64_t s;
    uint64_t next() {
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    uint32_t below(uint32_t n) { return n ? (uint32_t)(next() % n) : 0u; }
};

uint32_t nasty_word(Rng& r) {
    static const float fl[] = {0.0f, -0.0f, NAN, INFINITY, -INFINITY, 1e38f, -1e38f, 1e-45f, 3.5f, -7.25f};
    switch (r.below(8)) {
    case 0: return 0u;
    case 1: return 0xFFFFFFFFu;
    case 2: return (r.below(16) << 28) | r.below(1u << 28);       // any type tag, any index
    case 3: return (r.below(7) + 1u) << 28 | r.below(8);          // a valid tag with a small index
    case 4: return r.below(40);
    case 5: { float f = fl[r.below(10)]; uint32_t u; std::memcpy(&u, &f, 4); return u; }
    case 6: { float f = (float)((double)r.below(2000001) / 1000.0 - 1000.0); uint32_t u; std::memcpy(&u, &f, 4); return u; }
    default: return (uint32_t)r.next();
    }
}

// A lowered scene in heap arrays of exactly the advertised length.
struct Owned {
    vk_scene_desc d{};
    std::vector<vk_node> nodes; std::vector<vk_sphere> spheres; std::vector<uint32_t> sphere_mat; std::vector<vk_msphere> mspheres;
    std::vector<vk_rect> rects; std::vector<vk_box> boxes; std::vector<vk_xform> xforms; std::vector<vk_medium> media;
    std::vector<vk_ref> lights; std::vector<vk_material> materials; std::vector<vk_texture> textures; std::vector<uint8_t> texels;
    std::vector<vk_perlin> perlins;

    template <class T> static std::vector<T> exact(const T* p, size_t n) {
        std::vector<T> v;
        v.reserve(n); // capacity == size: the byte after the last element is poisoned
        v.assign(p, p + n);
        return v;
    }
    explicit Owned(const vk_scene_desc& s) : d(s) {
        nodes = exact(s.nodes, s.n_nodes); spheres = exact(s.spheres, s.n_spheres); sphere_mat = exact(s.sphere_mat, s.n_spheres);
        mspheres = exact(s.mspheres, s.n_mspheres); rects = exact(s.rects, s.n_rects); boxes = exact(s.boxes, s.n_boxes);
        xforms = exact(s.xforms, s.n_xforms); media = exact(s.media, s.n_media); lights = exact(s.lights, s.n_lights);
        materials = exact(s.materials, s.n_materials); textures = exact(s.textures, s.n_textures);
        texels = exact(s.texels, (size_t)s.n_texel_bytes); perlins = exact(s.perlins, s.n_perlins);
        bind();
    }
    template <class T> static const T* ptr(const std::vector<T>& v) { return v.empty() ? nullptr : v.data(); }
    void bind() {
        d.nodes = ptr(nodes); d.n_nodes = (uint32_t)nodes.size();
        d.spheres = ptr(spheres); d.sphere_mat = ptr(sphere_mat); d.n_spheres = (uint32_t)spheres.size();
        d.mspheres = ptr(mspheres); d.n_mspheres = (uint32_t)mspheres.size();
        d.rects = ptr(rects); d.n_rects = (uint32_t)rects.size();
        d.boxes = ptr(boxes); d.n_boxes = (uint32_t)boxes.size();
        d.xforms = ptr(xforms); d.n_xforms = (uint32_t)xforms.size();
        d.media = ptr(media); d.n_media = (uint32_t)media.size();
        d.lights = ptr(lights); d.n_lights = (uint32_t)lights.size();
        d.materials = ptr(materials); d.n_materials = (uint32_t)materials.size();
        d.textures = ptr(textures); d.n_textures = (uint32_t)textures.size();
        d.texels = ptr(texels); d.n_texel_bytes = texels.size();
        d.perlins = ptr(perlins); d.n_perlins = (uint32_t)perlins.size();
    }
    template <class T> static void shrink(std::vector<T>& v, Rng& r) {
        if (v.empty()) return;
        std::vector<T> w;
        const size_t n = r.below((uint32_t)v.size());
        w.reserve(n);
        w.assign(v.begin(), v.begin() + n);
        v.swap(w);
    }
    template <class T> static void scribble(std::vector<T>& v, Rng& r) {
        if (v.empty()) return;
        static_assert(sizeof(T) % 4 == 0, "records are made of 32-bit words");
        uint32_t* words = reinterpret_cast<uint32_t*>(v.data());
        const size_t n_words = v.size() * (sizeof(T) / 4);
        const uint32_t w = nasty_word(r);
        std::memcpy(words + r.below((uint32_t)n_words), &w, 4);
    }
    void corrupt(Rng& r) {
        switch (r.below(16)) {
        case 0: d.root = nasty_word(r); return;
        case 1: shrink(nodes, r); break;
        case 2: shrink(materials, r); break;
        case 3: shrink(textures, r); break;
        case 4: switch (r.below(8)) {
                case 0: shrink(spheres, r); sphere_mat.resize(spheres.size()); { std::vector<uint32_t> w; w.reserve(sphere_mat.size()); w.assign(sphere_mat.begin(), sphere_mat.end()); sphere_mat.swap(w); } break;
                case 1: shrink(rects, r); break;
                case 2: shrink(boxes, r); break;
                case 3: shrink(xforms, r); break;
                case 4: shrink(media, r); break;
                case 5: shrink(lights, r); break;
                case 6: shrink(perlins, r); break;
                default: shrink(texels, r); break;
                }
                break;
        case 5: scribble(nodes, r); break;
        case 6: scribble(nodes, r); break;
        case 7: scribble(spheres, r); break;
        case 8: scribble(sphere_mat, r); break;
        case 9: scribble(rects, r); break;
        case 10: scribble(boxes, r); break;
        case 11: scribble(xforms, r); break;
        case 12: scribble(media, r); break;
        case 13: scribble(lights, r); break;
        case 14: scribble(materials, r); break;
        default: scribble(textures, r); if (!mspheres.empty() && r.below(2)) scribble(mspheres, r); break;
        }
        bind();
    }
};

int fail(const char* what, const char* why) {
    std::fprintf(stderr, "sanitize_host: %s: %s\n", what, why);
    return 1;
}

} // namespace

int main(int argc, char** argv) {
    const char* assets = argc > 1 ? argv[1] : "assets";
    const char* tmp_ppm = argc > 2 ? argv[2] : "/tmp/sanitize_host.ppm";
    const int fuzz_cases = argc > 3 ? std::atoi(argv[3]) : 400;
    const uint64_t salt = argc > 4 ? std::strtoull(argv[4], nullptr, 0) : 0; // another stream of corruptions

    struct Job { const char* name; uint32_t param; bool fuzz; };
    const Job jobs[] = {{"cornell_box", 0, true},      {"cornell_smoke", 0, true},       {"final_scene", 0, true},   {"random_spheres_demo", 0, true},
                        {"bowser_demo", 0, true},      {"perlin_demo", 0, true},         {"balls_demo", 0, false},   {"api_surface_demo", 0, true},
                        {"random_spheres_cover", 0, false}, {"book1_cover", 0, false},   {"stress_spheres", 40, true}, {"furnace_demo", 0, false},
                        {"furnace_demo", 1, false},    {"furnace_demo", 2, false},       {"furnace_demo", 3, true}};
    unsigned long long accepted = 0, refused = 0, cameras = 0;
    for (const Job& job : jobs) {
        for (uint64_t seed = 1; seed <= 2; ++seed) {
            vkh_scene* sc = nullptr;
            if (vkh_scene_build(job.name, seed, assets, job.param, &sc) != VK_OK) return fail(job.name, vkh_last_error());
            const vk_scene_desc* d = vkh_scene_desc(sc);
            vk_scene_info info;
            char why[256];
            if (vk_scene_check(d, &info, why, sizeof why) != VK_OK) return fail(job.name, why);
            vk_camera cam;
            while (vkh_scene_next_camera(sc, &cam)) ++cameras; // FixedCamera: 1; RotatingCamera: 671 / 721
            if (seed == 1 && job.fuzz) {
                Rng r{0xC0FFEEull * (uint64_t)(&job - jobs + 1) + salt * 0xD1B54A32D192ED03ull}; // (not a small multiple of the generator's increment)
                // the 15 MB of Bowser textures are copied per case: fewer cases there
                const int cases = d->n_texel_bytes > (4u << 20) ? fuzz_cases / 8 : fuzz_cases;
                for (int c = 0; c < cases; ++c) {
                    Owned o(*d);
                    for (uint32_t k = 0, n = 1 + r.below(3); k < n; ++k) o.corrupt(r);
                    (vk_scene_check(&o.d, &info, why, sizeof why) == VK_OK ? accepted : refused)++;
                    if (c % 7 == 0) vk_scene_check(&o.d, nullptr, nullptr, 0); // the optional outputs
                }
            }
            vkh_scene_free(sc);
        }
    }
    if (vk_scene_check(nullptr, nullptr, nullptr, 0) == VK_OK) return fail("vk_scene_check(NULL)", "accepted");

    // the output side of the frame loop: to_color of awkward values, the P3 writer, the file name rule
    const uint32_t W = 5, H = 3;
    std::vector<float> frame(W * H * 3);
    const float awkward[] = {0.0f, 1.0f, 0.25f, 2.0f, -1.0f, NAN, INFINITY, -INFINITY, 1e-30f, 0.999f};
    for (size_t i = 0; i < frame.size(); ++i) frame[i] = awkward[i % 10];
    std::vector<uint8_t> rgb8(frame.size());
    vkh_frame_to_rgb8(frame.data(), W, H, rgb8.data());
    if (vkh_write_ppm(tmp_ppm, rgb8.data(), W, H) != VK_OK) return fail("vkh_write_ppm", vkh_last_error());
    if (vkh_write_ppm("/nonexistent-dir/x.ppm", rgb8.data(), W, H) == VK_OK) return fail("vkh_write_ppm", "wrote into a missing directory");
    char name[16];
    if (vkh_frame_filename("", 7, name, sizeof name) < 0 || std::strcmp(name, "output_0007.ppm")) return fail("vkh_frame_filename", name);
    if (vkh_frame_filename("some/dir", 7, name, sizeof name) >= 0) return fail("vkh_frame_filename", "did not notice the short buffer");

    // the PNG decoder on a truncated file and on a file that is not a PNG
    uint32_t w = 0, h = 0;
    if (vkh_decode_png(tmp_ppm, nullptr, 0, &w, &h) >= 0) return fail("vkh_decode_png", "decoded a P3 file");
    {
        const std::string src = std::string(assets) + "/earthmap.png", cut = std::string(tmp_ppm) + ".png";
        FILE* in = std::fopen(src.c_str(), "rb");
        if (!in) return fail("fopen", src.c_str());
        std::vector<uint8_t> bytes(20000);
        bytes.resize(std::fread(bytes.data(), 1, bytes.size(), in));
        std::fclose(in);
        FILE* out = std::fopen(cut.c_str(), "wb");
        if (!out) return fail("fopen", cut.c_str());
        std::fwrite(bytes.data(), 1, bytes.size(), out);
        std::fclose(out);
        if (vkh_decode_png(cut.c_str(), nullptr, 0, &w, &h) >= 0) {
            std::vector<uint8_t> px((size_t)w * h * 3);
            if (vkh_decode_png(cut.c_str(), px.data(), px.size(), &w, &h) >= 0) return fail("vkh_decode_png", "decoded a truncated PNG");
        }
        std::remove(cut.c_str());
    }
    std::printf("sanitize_host ok: %llu cameras, %llu corrupted scenes refused, %llu accepted and planned\n", cameras, refused, accepted);
    return 0;
}
