// vecchio.cpp -- host front end: constructors, bounding boxes, BVH build, camera, PNG decode,
// Perlin tables and lower().  See vecchio.hpp for the scope statement.
#include "vecchio.hpp"

#include <zlib.h>

#include <cstdio>
#include <fstream>

namespace vecchio {

static HostRng g_rng(1);
HostRng& thread_rng() { return g_rng; }
void seed_thread_rng(uint64_t seed) { g_rng.reseed(seed); }
std::string g_assets_dir = "assets";

// f32::min / f32::max: the non-NaN operand wins (src/util.rs:6-12)
static inline float fmin_(float a, float b) { return std::fmin(a, b); }
static inline float fmax_(float a, float b) { return std::fmax(a, b); }

// ------------------------------------------------------------------ Camera (src/main.rs:71-109)
Camera Camera::make(Vec3 lookfrom, Vec3 lookat, Vec3 vup, float vfov, float aspect_ratio,
                    float aperture, float focus_dist, float time0, float time1) {
    float theta = to_radians(vfov);
    float h = std::tan(theta / 2.0f);
    float viewport_height = h * 2.0f;
    float viewport_width = aspect_ratio * viewport_height;

    Vec3 w = (lookfrom - lookat).unit_vector();
    Vec3 u = vup.cross(w).unit_vector();
    Vec3 v = w.cross(u);

    Camera c;
    c.origin = lookfrom;
    c.horizontal = u * viewport_width * focus_dist;
    c.vertical = v * viewport_height * focus_dist;
    c.lower_left_corner = c.origin - c.horizontal / 2.0f - c.vertical / 2.0f - w * focus_dist;
    c.lens_radius = aperture / 2.0f;
    c.u = u;
    c.v = v;
    c.w = w;
    c.time0 = time0;
    c.time1 = time1;
    return c;
}

vk_camera Camera::lower() const {
    vk_camera k;
    auto put = [](float* d, Vec3 s) { d[0] = s.x; d[1] = s.y; d[2] = s.z; };
    put(k.origin, origin);
    put(k.lower_left_corner, lower_left_corner);
    put(k.horizontal, horizontal);
    put(k.vertical, vertical);
    put(k.u, u);
    put(k.v, v);
    put(k.w, w);
    k.lens_radius = lens_radius;
    k.time0 = time0;
    k.time1 = time1;
    return k;
}

// ------------------------------------------------------------------ AxisBB (src/accel.rs:37-49)
AxisBB AxisBB::surrounding_box(AxisBB b1, AxisBB b2) {
    Vec3 small(fmin_(b1.min.x, b2.min.x), fmin_(b1.min.y, b2.min.y), fmin_(b1.min.z, b2.min.z));
    Vec3 big(fmax_(b1.max.x, b2.max.x), fmax_(b1.max.y, b2.max.y), fmax_(b1.max.z, b2.max.z));
    return AxisBB(small, big);
}

// ------------------------------------------------------------------ bounding boxes
std::optional<AxisBB> Sphere::bounding_box(float, float) const { // src/hittable.rs:97-102
    return AxisBB(center - Vec3::new_const(radius), center + Vec3::new_const(radius));
}
Vec3 MovingSphere::center(float time) const { // src/hittable.rs:147-150
    return center0 + (center1 - center0) * ((time - time0) / (time1 - time0));
}
std::optional<AxisBB> MovingSphere::bounding_box(float, float) const { // src/hittable.rs:186-196
    AxisBB bb1(center(time0) - Vec3::new_const(radius), center(time0) + Vec3::new_const(radius));
    AxisBB bb2(center(time1) - Vec3::new_const(radius), center(time1) + Vec3::new_const(radius));
    return AxisBB::surrounding_box(bb1, bb2);
}
std::optional<AxisBB> Rect::bounding_box(float, float) const { // src/hittable.rs:258-269
    Vec3 v1 = Vec3::zero(), v2 = Vec3::zero();
    v1[axis0] = c0;
    v1[axis1] = d0;
    v1[axis2] = k - 0.0001f;
    v2[axis0] = c1;
    v2[axis1] = d1;
    v2[axis2] = k + 0.0001f;
    return AxisBB(v1, v2);
}
Boxy::Boxy(Vec3 p0, Vec3 p1, Arc<MaterialSS> m) : box_min(p0), box_max(p1), mat(std::move(m)) {
    // assert!(p0.x < p1.x) ... src/hittable.rs:322-324
    if (!(p0.x < p1.x) || !(p0.y < p1.y) || !(p0.z < p1.z)) throw std::invalid_argument("Boxy::new: p0 must be < p1");
}
std::optional<AxisBB> Translate::bounding_box(float t0, float t1) const { // src/hittable.rs:525-531
    auto bb = ptr->bounding_box(t0, t1);
    if (!bb) return std::nullopt;
    return AxisBB(bb->min + offset, bb->max + offset);
}
RotateAxis::RotateAxis(Arc<HittableSS> p, float angle, uint32_t kind_) : ptr(std::move(p)), kind(kind_) {
    // src/hittable.rs:542-575 (Y), :639-672 (X), :728-761 (Z)
    sin_theta = std::sin(to_radians(angle));
    cos_theta = std::cos(to_radians(angle));
    auto bbox = ptr->bounding_box(0.0f, 1.0f);
    if (!bbox) throw std::invalid_argument("Rotate::new: child has no bounding box");
    Vec3 mn = Vec3::new_const(INFINITY), mx = Vec3::new_const(-INFINITY);
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j)
            for (int k = 0; k < 2; ++k) {
                float x = i == 1 ? bbox->max.x : bbox->min.x;
                float y = j == 1 ? bbox->max.y : bbox->min.y;
                float z = k == 1 ? bbox->max.z : bbox->min.z;
                Vec3 tester;
                if (kind == VK_X_ROTATE_Y) {
                    tester = Vec3(cos_theta * x + sin_theta * z, y, -sin_theta * x + cos_theta * z);
                } else if (kind == VK_X_ROTATE_X) {
                    tester = Vec3(x, cos_theta * y - sin_theta * z, sin_theta * y + cos_theta * z);
                } else {
                    tester = Vec3(cos_theta * x - sin_theta * y, sin_theta * x + cos_theta * y, z);
                }
                for (size_t c = 0; c < 3; ++c) {
                    mn[c] = fmin_(mn[c], tester[c]);
                    mx[c] = fmax_(mx[c], tester[c]);
                }
            }
    bb = AxisBB(mn, mx);
}

// ------------------------------------------------------------------ BVHNode::new (src/accel.rs:98-136)
static AxisBB bb_of(const Arc<HittableSS>& h) {
    auto bb = h->bounding_box(0.0f, 0.0f);
    if (!bb) throw std::invalid_argument("BVHNode::new: object without bounding box");
    return *bb;
}
Arc<BVHNode> BVHNode::make(Arc<HittableSS>* objects, size_t n) {
    if (n == 0) throw std::invalid_argument("BVHNode::new: empty object list");
    auto& rng = thread_rng();
    size_t axis = rng.gen_range_u32(0, 3);
    auto node = std::make_shared<BVHNode>();
    if (n == 1) {
        node->left = objects[0];
        node->right = objects[0];
        node->bb = AxisBB::surrounding_box(bb_of(objects[0]), bb_of(objects[0]));
    } else if (n == 2) {
        AxisBB a_bb = bb_of(objects[0]), b_bb = bb_of(objects[1]);
        size_t i1 = 0, i2 = 1;
        if (a_bb.min[axis] < b_bb.min[axis]) { // the larger min goes LEFT (:111-115)
            i1 = 1;
            i2 = 0;
        }
        node->left = objects[i1];
        node->right = objects[i2];
        node->bb = AxisBB::surrounding_box(bb_of(objects[i1]), bb_of(objects[i2]));
    } else {
        // objects.sort_by(bb.min[axis]) -- Rust's sort_by is stable; keys are pure, so they are
        // computed once per object instead of once per comparison.
        std::vector<std::pair<float, Arc<HittableSS>>> keyed(n);
        for (size_t i = 0; i < n; ++i) {
            float key = bb_of(objects[i]).min[axis];
            if (std::isnan(key)) throw std::invalid_argument("BVHNode::new: NaN bounding box"); // partial_cmp().unwrap()
            keyed[i] = {key, std::move(objects[i])};
        }
        std::stable_sort(keyed.begin(), keyed.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
        for (size_t i = 0; i < n; ++i) objects[i] = std::move(keyed[i].second);
        keyed.clear();
        keyed.shrink_to_fit();
        size_t mid = n / 2;
        auto l = BVHNode::make(objects, mid);
        auto r = BVHNode::make(objects + mid, n - mid);
        node->bb = AxisBB::surrounding_box(l->bb, r->bb);
        node->left = std::move(l);
        node->right = std::move(r);
    }
    return node;
}

// ------------------------------------------------------------------ Perlin::new (src/material.rs:357-377)
Perlin::Perlin() {
    auto& rng = thread_rng();
    for (auto& v : random_data) v = Vec3::random_range(-1.0f, 1.0f).unit_vector();
    for (size_t i = 0; i < NUM_POINTS; ++i) perm_x[i] = perm_y[i] = perm_z[i] = i;
    rng.shuffle(perm_x.data(), NUM_POINTS);
    rng.shuffle(perm_y.data(), NUM_POINTS);
    rng.shuffle(perm_z.data(), NUM_POINTS);
}

// ------------------------------------------------------------------ PNG decode (src/material.rs:269-279)
// The reference uses the `png` crate; PNG is lossless, so any decoder yields the same bytes.
// 8-bit, non-interlaced RGB only -- which is what BPP = 3 (src/material.rs:267) assumes.
static uint32_t be32(const uint8_t* p) { return (uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | p[3]; }
static void decode_png_rgb8(const std::string& path, size_t& w, size_t& h, std::vector<uint8_t>& out) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error("ImageTexture::new: cannot open " + path);
    std::vector<uint8_t> file((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    if (file.size() < 8 || std::memcmp(file.data(), sig, 8) != 0) throw std::runtime_error("not a PNG: " + path);
    std::vector<uint8_t> idat;
    size_t pos = 8;
    bool have_ihdr = false;
    while (pos + 12 <= file.size()) {
        uint32_t len = be32(&file[pos]);
        const uint8_t* type = &file[pos + 4];
        const uint8_t* data = &file[pos + 8];
        if (pos + 12 + (size_t)len > file.size()) throw std::runtime_error("truncated PNG: " + path);
        if (!std::memcmp(type, "IHDR", 4)) {
            w = be32(data);
            h = be32(data + 4);
            if (data[8] != 8 || data[9] != 2 || data[12] != 0)
                throw std::runtime_error("PNG must be 8-bit non-interlaced RGB: " + path);
            have_ihdr = true;
        } else if (!std::memcmp(type, "IDAT", 4)) {
            idat.insert(idat.end(), data, data + len);
        } else if (!std::memcmp(type, "IEND", 4)) {
            break;
        }
        pos += 12 + (size_t)len;
    }
    if (!have_ihdr) throw std::runtime_error("PNG without IHDR: " + path);
    const size_t bpp = 3, stride = w * bpp;
    std::vector<uint8_t> raw((stride + 1) * h);
    uLongf raw_len = (uLongf)raw.size();
    if (uncompress(raw.data(), &raw_len, idat.data(), (uLong)idat.size()) != Z_OK || raw_len != raw.size())
        throw std::runtime_error("PNG inflate failed: " + path);
    out.assign(stride * h, 0);
    for (size_t y = 0; y < h; ++y) {
        const uint8_t ft = raw[y * (stride + 1)];
        const uint8_t* src = &raw[y * (stride + 1) + 1];
        uint8_t* dst = &out[y * stride];
        const uint8_t* up = y ? &out[(y - 1) * stride] : nullptr;
        for (size_t i = 0; i < stride; ++i) {
            int a = i >= bpp ? dst[i - bpp] : 0;
            int b = up ? up[i] : 0;
            int c = (up && i >= bpp) ? up[i - bpp] : 0;
            int pred = 0;
            switch (ft) {
            case 0: pred = 0; break;
            case 1: pred = a; break;
            case 2: pred = b; break;
            case 3: pred = (a + b) >> 1; break;
            case 4: {
                int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
                pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
                break;
            }
            default: throw std::runtime_error("bad PNG filter: " + path);
            }
            dst[i] = (uint8_t)(src[i] + pred);
        }
    }
}
static bool file_exists(const std::string& p) { return (bool)std::ifstream(p, std::ios::binary); }
ImageTexture::ImageTexture(const std::string& path) {
    std::string p = path;
    if (!file_exists(p)) { // "assets/earthmap.png" relative to the crate root in the reference
        size_t slash = path.find_last_of('/');
        p = g_assets_dir + "/" + (slash == std::string::npos ? path : path.substr(slash + 1));
    }
    decode_png_rgb8(p, width, height, buf);
}

// ------------------------------------------------------------------ lower(): textures
uint32_t Lowering::texture(const Arc<TextureSS>& t) {
    auto it = memo_t.find(t.get());
    if (it != memo_t.end()) return it->second;
    uint32_t idx = t->lower(*this);
    memo_t[t.get()] = idx;
    return idx;
}
uint32_t SolidColor::lower(Lowering& L) const {
    vk_texture t{};
    t.type = VK_TEX_SOLID;
    t.rgb[0] = color_value.x;
    t.rgb[1] = color_value.y;
    t.rgb[2] = color_value.z;
    L.textures.push_back(t);
    return (uint32_t)L.textures.size() - 1;
}
uint32_t Checker::lower(Lowering& L) const {
    uint32_t o = L.texture(odd), e = L.texture(even);
    vk_texture t{};
    t.type = VK_TEX_CHECKER;
    t.checker.odd = o;
    t.checker.even = e;
    L.textures.push_back(t);
    return (uint32_t)L.textures.size() - 1;
}
uint32_t ImageTexture::lower(Lowering& L) const {
    if (L.texels.size() + buf.size() > 0xFFFFFFFFull) throw LowerError("ImageTexture: texel pool exceeds 4 GiB");
    vk_texture t{};
    t.type = VK_TEX_IMAGE;
    t.image.texel_offset = (uint32_t)L.texels.size();
    t.image.width = (uint32_t)width;
    t.image.height = (uint32_t)height;
    L.texels.insert(L.texels.end(), buf.begin(), buf.end());
    L.textures.push_back(t);
    return (uint32_t)L.textures.size() - 1;
}
uint32_t NoiseTexture::lower(Lowering& L) const {
    vk_perlin p{};
    for (size_t i = 0; i < Perlin::NUM_POINTS; ++i) {
        p.ranvec[i][0] = noise.random_data[i].x;
        p.ranvec[i][1] = noise.random_data[i].y;
        p.ranvec[i][2] = noise.random_data[i].z;
        p.perm_x[i] = (uint8_t)noise.perm_x[i];
        p.perm_y[i] = (uint8_t)noise.perm_y[i];
        p.perm_z[i] = (uint8_t)noise.perm_z[i];
    }
    L.perlins.push_back(p);
    vk_texture t{};
    t.type = VK_TEX_NOISE;
    t.noise.perlin = (uint32_t)L.perlins.size() - 1;
    t.noise.scale = scale;
    L.textures.push_back(t);
    return (uint32_t)L.textures.size() - 1;
}

// ------------------------------------------------------------------ lower(): materials
uint32_t Lowering::material(const Arc<MaterialSS>& m) {
    auto it = memo_m.find(m.get());
    if (it != memo_m.end()) return it->second;
    uint32_t idx = m->lower(*this);
    memo_m[m.get()] = idx;
    return idx;
}
static uint32_t push_mat(Lowering& L, uint32_t type, uint32_t tex, float param, uint32_t aux) {
    L.materials.push_back(vk_material{type, tex, param, aux});
    return (uint32_t)L.materials.size() - 1;
}
uint32_t Lambertian::lower(Lowering& L) const { return push_mat(L, VK_M_LAMBERTIAN, L.texture(albedo), 0.0f, 0); }
uint32_t Metal::lower(Lowering& L) const { return push_mat(L, VK_M_METAL, L.texture(albedo), fuzz, 0); }
uint32_t Dielectric::lower(Lowering& L) const { return push_mat(L, VK_M_DIELECTRIC, 0, ref_idx, 0); }
uint32_t DiffuseLight::lower(Lowering& L) const { return push_mat(L, VK_M_DIFFUSE_LIGHT, L.texture(emit), 0.0f, 0); }
uint32_t Isotropic::lower(Lowering& L) const { return push_mat(L, VK_M_ISOTROPIC, L.texture(albedo), 0.0f, 0); }
uint32_t SpecDiffuse::lower(Lowering& L) const {
    if (dynamic_cast<const SpecDiffuse*>(specular.get()) || dynamic_cast<const SpecDiffuse*>(diffuse.get()))
        throw LowerError("SpecDiffuse: nested SpecDiffuse is not supported on the GPU path");
    uint32_t s = L.material(specular), d = L.material(diffuse);
    return push_mat(L, VK_M_SPECDIFFUSE, s, pct, d);
}

// ------------------------------------------------------------------ lower(): hittables
vk_ref Lowering::hittable(const Arc<HittableSS>& h) {
    auto it = memo_h.find(h.get());
    if (it != memo_h.end()) return it->second;
    vk_ref r = h->lower(*this);
    memo_h[h.get()] = r;
    return r;
}
vk_ref Sphere::lower(Lowering& L) const {
    uint32_t m = L.material(material);
    L.spheres.push_back(vk_sphere{{center.x, center.y, center.z}, radius});
    L.sphere_mat.push_back(m);
    return VK_REF(VK_T_SPHERE, L.spheres.size() - 1);
}
vk_ref MovingSphere::lower(Lowering& L) const {
    vk_msphere s{};
    s.center0[0] = center0.x; s.center0[1] = center0.y; s.center0[2] = center0.z;
    s.center1[0] = center1.x; s.center1[1] = center1.y; s.center1[2] = center1.z;
    s.radius = radius;
    s.time0 = time0;
    s.time1 = time1;
    s.mat = L.material(material);
    L.mspheres.push_back(s);
    return VK_REF(VK_T_MSPHERE, L.mspheres.size() - 1);
}
vk_ref Rect::lower_flipped(Lowering& L, bool flip) const {
    auto& memo = flip ? L.memo_h_flip : L.memo_h;
    auto it = memo.find(this);
    if (it != memo.end()) return it->second;
    vk_rect r{};
    r.c0 = c0; r.c1 = c1; r.d0 = d0; r.d1 = d1; r.k = k;
    r.axes = (uint32_t)axis0 | ((uint32_t)axis1 << 2) | ((uint32_t)axis2 << 4) | (flip ? VK_RECT_FLIP : 0u);
    r.mat = L.material(mat);
    L.rects.push_back(r);
    vk_ref ref = VK_REF(VK_T_RECT, L.rects.size() - 1);
    memo[this] = ref;
    return ref;
}
vk_ref Rect::lower(Lowering& L) const { return lower_flipped(L, false); }
vk_ref FlipFace::lower(Lowering& L) const {
    // FlipFace(Rect) -- every use in src/scene.rs and Boxy::new -- folds into the rect record.
    if (auto* r = dynamic_cast<const Rect*>(ptr.get())) return r->lower_flipped(L, true);
    vk_xform x{};
    x.kind = VK_X_FLIP;
    x.child = L.hittable(ptr);
    L.xforms.push_back(x);
    return VK_REF(VK_T_XFORM, L.xforms.size() - 1);
}
vk_ref Boxy::lower(Lowering& L) const {
    vk_box b{};
    b.box_min[0] = box_min.x; b.box_min[1] = box_min.y; b.box_min[2] = box_min.z;
    b.box_max[0] = box_max.x; b.box_max[1] = box_max.y; b.box_max[2] = box_max.z;
    b.mat = L.material(mat);
    L.boxes.push_back(b);
    return VK_REF(VK_T_BOX, L.boxes.size() - 1);
}
vk_ref ConstantMedium::lower(Lowering& L) const {
    vk_medium m{};
    m.boundary = L.hittable(boundary);
    m.neg_inv_density = neg_inv_density;
    m.mat = L.material(phase_function);
    L.media.push_back(m);
    return VK_REF(VK_T_MEDIUM, L.media.size() - 1);
}
vk_ref Translate::lower(Lowering& L) const {
    vk_xform x{};
    x.kind = VK_X_TRANSLATE;
    x.child = L.hittable(ptr);
    x.a = offset.x; x.b = offset.y; x.c = offset.z;
    L.xforms.push_back(x);
    return VK_REF(VK_T_XFORM, L.xforms.size() - 1);
}
vk_ref RotateAxis::lower(Lowering& L) const {
    vk_xform x{};
    x.kind = kind;
    x.child = L.hittable(ptr);
    x.a = sin_theta;
    x.b = cos_theta;
    L.xforms.push_back(x);
    return VK_REF(VK_T_XFORM, L.xforms.size() - 1);
}
vk_ref BVHNode::lower(Lowering& L) const {
    // depth-first (pre-order) numbering: a node's left subtree follows it immediately
    uint32_t idx = (uint32_t)L.nodes.size();
    L.nodes.push_back(vk_node{});
    vk_ref l = L.hittable(left);
    vk_ref r = (right.get() == left.get()) ? l : L.hittable(right);
    vk_node n{};
    n.bb_min[0] = bb.min.x; n.bb_min[1] = bb.min.y; n.bb_min[2] = bb.min.z;
    n.bb_max[0] = bb.max.x; n.bb_max[1] = bb.max.y; n.bb_max[2] = bb.max.z;
    n.left = l;
    n.right = r;
    L.nodes[idx] = n;
    return VK_REF(VK_T_NODE, idx);
}

vk_scene_desc Lowering::desc() const {
    vk_scene_desc d{};
    d.api_version = VK_API_VERSION;
    d.root = root;
    d.nodes = nodes.data();         d.n_nodes = (uint32_t)nodes.size();
    d.spheres = spheres.data();     d.sphere_mat = sphere_mat.data(); d.n_spheres = (uint32_t)spheres.size();
    d.mspheres = mspheres.data();   d.n_mspheres = (uint32_t)mspheres.size();
    d.rects = rects.data();         d.n_rects = (uint32_t)rects.size();
    d.boxes = boxes.data();         d.n_boxes = (uint32_t)boxes.size();
    d.xforms = xforms.data();       d.n_xforms = (uint32_t)xforms.size();
    d.media = media.data();         d.n_media = (uint32_t)media.size();
    d.lights = lights.data();       d.n_lights = (uint32_t)lights.size();
    d.materials = materials.data(); d.n_materials = (uint32_t)materials.size();
    d.textures = textures.data();   d.n_textures = (uint32_t)textures.size();
    d.texels = texels.data();       d.n_texel_bytes = texels.size();
    d.perlins = perlins.data();     d.n_perlins = (uint32_t)perlins.size();
    return d;
}

// main(): `BVHNode::new(&mut config.world[..])`, `Arc::new(config.lights)` (src/main.rs:168-169)
std::unique_ptr<LoweredScene> lower_scene(SceneConfig&& cfg) {
    auto out = std::make_unique<LoweredScene>();
    Arc<HittableSS> world_bvh = BVHNode::make(cfg.world);
    out->low.root = out->low.hittable(world_bvh);
    for (auto& l : cfg.lights) out->low.lights.push_back(out->low.hittable(l));
    out->cam_iter = std::move(cfg.cam_iter);
    out->aspect_ratio = cfg.aspect_ratio;
    return out;
}

} // namespace vecchio
