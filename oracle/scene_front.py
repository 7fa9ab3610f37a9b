"""oracle/scene_front.py -- the oracle's OWN scene front end.  TEST INFRASTRUCTURE ONLY (see oracle/oracle.h).

The C++ oracle (oracle.cpp) rebuilds its object graph from the product's flattened scene, so a wrong constant in
vecchio_b200/host/scene.cpp or a wrong field in one of the `lower()` methods would be common to both sides of every
oracle-vs-GPU test.  This module is a second, independent path to the same data: the reference's scene builders
restated in plain Python straight from the Rust --

    src/scene.rs:167-284   random_spheres_demo        src/scene.rs:630-730  cornell_box
    src/scene.rs:732-874   final_scene                SURVEY 8(d)           Cornell smoke (authored, book 2)
    src/scene.rs:93-165    balls_demo                 src/scene.rs:286-337  perlin_demo
    src/scene.rs:340-628   Bowser::new, bowser_demo (RotateX / RotateZ::new: src/hittable.rs:639-674, 728-763)
    src/accel.rs:36-50, 98-136   AxisBB::surrounding_box, BVHNode::new (random axis, stable sort, len/2 split)
    src/hittable.rs:97-102, 186-196, 258-269, 367-369, 495-497, 525-531, 542-574   bounding_box of every type
    src/material.rs:357-377      Perlin::new          src/main.rs:71-109    Camera::new
    src/vec3.rs:68-82            Vec3::random / random_range

-- with every f32 operation done in numpy float32 (no contraction), and `unlower()`, which reads the product's
`vk_scene_desc` back into the same nested form.  tests/test_scene_front.py compares the two trees node by node:
every constant, every random draw (same seeded stream, same draw ORDER as the reference's code), every BVH split.

Stand-in for the reference's unseeded rand::thread_rng(): splitmix64 seeded as documented in
vecchio_b200/host/vecchio.hpp (the seed -> stream rule is part of the scene API's contract, restated here)."""
import math
import os

import numpy as np

F = np.float32
M64 = (1 << 64) - 1


class Rng:
    def __init__(self, seed):
        self.s = (seed * 0x9E3779B97F4A7C15 + 0xD1B54A32D192ED03) & M64
        for _ in range(4):
            self.next_u64()

    def next_u64(self):
        self.s = (self.s + 0x9E3779B97F4A7C15) & M64
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
        return z ^ (z >> 31)

    def next_u32(self):
        return self.next_u64() >> 32

    def gen_f32(self):  # rand 0.7.3 gen::<f32>(): 24 bits, [0, 1)
        return F(self.next_u32() >> 8) * F(1.0 / 16777216.0)

    def gen_range(self, low, high):  # rand 0.7.3 UniformFloat: a value in [1, 2) * scale + (low - scale), redrawn if >= high
        low, high = F(low), F(high)
        scale = high - low
        offset = low - scale
        while True:
            v12 = np.uint32(0x3F800000 | (self.next_u32() >> 9)).view(F)
            res = v12 * scale + offset
            if res < high:
                return res

    def gen_range_u32(self, low, high):
        return low + ((self.next_u32() * (high - low)) >> 32)

    def shuffle(self, a):  # SliceRandom::shuffle (Fisher-Yates from the top)
        for i in range(len(a), 1, -1):
            j = self.gen_range_u32(0, i)
            a[i - 1], a[j] = a[j], a[i - 1]


def v3(x, y, z):
    return np.array([x, y, z], dtype=F)


def vrandom(rng):  # Vec3::random src/vec3.rs:68-72
    return v3(rng.gen_f32(), rng.gen_f32(), rng.gen_f32())


def vrandom_range(rng, lo, hi):  # :74-82
    return v3(rng.gen_range(lo, hi), rng.gen_range(lo, hi), rng.gen_range(lo, hi))


def unit(v):  # three divisions by sqrt(len2), src/vec3.rs:39-42
    n = np.sqrt(F(v[0] * v[0] + v[1] * v[1]) + v[2] * v[2])
    return v3(v[0] / n, v[1] / n, v[2] / n)


def cross(a, b):
    return v3(a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0])


# ------------------------------------------------------------------------------------------------ textures / materials
def solid(r, g, b):
    return {"kind": "solid", "rgb": v3(r, g, b)}


def solid_v(v):
    return {"kind": "solid", "rgb": np.asarray(v, dtype=F)}


def lambertian(tex):
    return {"kind": "lambertian", "tex": tex}


def metal(tex, fuzz):
    return {"kind": "metal", "tex": tex, "fuzz": F(fuzz)}


def dielectric(ior):
    return {"kind": "dielectric", "ior": F(ior)}


def diffuse_light(tex):
    return {"kind": "diffuse_light", "tex": tex}


def isotropic(tex):
    return {"kind": "isotropic", "tex": tex}


def perlin_new(rng):  # src/material.rs:357-377: 256 unit vectors, then the three shuffles
    vecs = np.stack([unit(vrandom_range(rng, -1.0, 1.0)) for _ in range(256)])
    perms = []
    for _ in range(3):
        p = list(range(256))
        rng.shuffle(p)
        perms.append(np.array(p, dtype=np.uint8))
    return {"ranvec": vecs, "perm_x": perms[0], "perm_y": perms[1], "perm_z": perms[2]}


def noise_texture(rng, scale):  # NoiseTexture::new -> Perlin::new (:422-427)
    return {"kind": "noise", "scale": F(scale), "perlin": perlin_new(rng)}


def image_texture(path, assets_dir):  # ImageTexture::new :269-279 -- decoded here with PIL, the product uses its own inflate
    from PIL import Image
    im = Image.open(os.path.join(assets_dir, os.path.basename(path))).convert("RGB")
    a = np.asarray(im, dtype=np.uint8)
    return {"kind": "image", "width": a.shape[1], "height": a.shape[0], "texels": a.reshape(-1).copy()}


# ------------------------------------------------------------------------------------------------ hittables
def sphere(c, r, mat):
    return {"kind": "sphere", "center": np.asarray(c, dtype=F), "radius": F(r), "mat": mat}


def moving_sphere(c0, c1, t0, t1, r, mat):
    return {"kind": "msphere", "center0": np.asarray(c0, dtype=F), "center1": np.asarray(c1, dtype=F), "time0": F(t0), "time1": F(t1),
            "radius": F(r), "mat": mat}


def rect(kind, c0, c1, d0, d1, k, mat, flip=False):  # Rect::XYRect / XZRect / YZRect :214-226
    axes = {"xy": (0, 1, 2), "xz": (0, 2, 1), "yz": (1, 2, 0)}[kind]
    return {"kind": "rect", "c0": F(c0), "c1": F(c1), "d0": F(d0), "d1": F(d1), "k": F(k), "axes": axes, "flip": flip, "mat": mat}


def flip_face(h):  # FlipFace::new(Rect) -- every use in scene.rs: normalised to a flag on the rect, as the lowering stores it
    assert h["kind"] == "rect"
    return dict(h, flip=True)


def boxy(p0, p1, mat):
    p0, p1 = np.asarray(p0, dtype=F), np.asarray(p1, dtype=F)
    assert np.all(p0 < p1)  # the three assert!s of Boxy::new :322-324
    return {"kind": "box", "min": p0, "max": p1, "mat": mat}


def translate(h, offset):
    return {"kind": "translate", "offset": np.asarray(offset, dtype=F), "child": h}


def _rotate(kind, h, angle, turn):  # RotateX / RotateY / RotateZ::new (:639-674, :542-574, :728-763): sin, cos, the rotated box
    rad = F(angle) * F(math.pi / 180.0)  # f32::to_radians
    s, c = F(np.sin(rad)), F(np.cos(rad))
    bmin, bmax = bounding_box(h)
    mn, mx = np.full(3, np.inf, dtype=F), np.full(3, -np.inf, dtype=F)
    for i in range(2):
        for j in range(2):
            for k in range(2):
                x = bmax[0] if i == 1 else bmin[0]
                y = bmax[1] if j == 1 else bmin[1]
                z = bmax[2] if k == 1 else bmin[2]
                t = turn(s, c, x, y, z)
                mn, mx = np.minimum(mn, t), np.maximum(mx, t)
    return {"kind": kind, "sin": s, "cos": c, "child": h, "bb": (mn, mx)}


def rotate_y(h, angle):  # :542-574
    return _rotate("rotate_y", h, angle, lambda s, c, x, y, z: v3(F(c * x) + F(s * z), y, F(-s * x) + F(c * z)))


def rotate_x(h, angle):  # :639-674
    return _rotate("rotate_x", h, angle, lambda s, c, x, y, z: v3(x, F(c * y) - F(s * z), F(s * y) + F(c * z)))


def rotate_z(h, angle):  # :728-763
    return _rotate("rotate_z", h, angle, lambda s, c, x, y, z: v3(F(c * x) - F(s * y), F(s * x) + F(c * y), z))


def constant_medium(boundary, density, tex):  # ConstantMedium::new :443-451
    return {"kind": "medium", "boundary": boundary, "neg_inv_density": F(-1.0) / F(density), "mat": isotropic(tex)}


def surrounding(a, b):
    return np.minimum(a[0], b[0]), np.maximum(a[1], b[1])


def bounding_box(h):
    k = h["kind"]
    if k == "sphere":
        r = np.full(3, h["radius"], dtype=F)
        return h["center"] - r, h["center"] + r
    if k == "msphere":  # ignores its arguments: the union of the boxes at time0 and time1 (:186-196)
        def center(t):
            return h["center0"] + (h["center1"] - h["center0"]) * F((t - h["time0"]) / (h["time1"] - h["time0"]))
        r = np.full(3, h["radius"], dtype=F)
        return surrounding((center(h["time0"]) - r, center(h["time0"]) + r), (center(h["time1"]) - r, center(h["time1"]) + r))
    if k == "rect":  # padded by 0.0001 on the plane's axis (:258-269)
        a0, a1, a2 = h["axes"]
        lo, hi = np.zeros(3, dtype=F), np.zeros(3, dtype=F)
        lo[a0], lo[a1], lo[a2] = h["c0"], h["d0"], h["k"] - F(0.0001)
        hi[a0], hi[a1], hi[a2] = h["c1"], h["d1"], h["k"] + F(0.0001)
        return lo, hi
    if k == "box":
        return h["min"], h["max"]
    if k == "translate":
        lo, hi = bounding_box(h["child"])
        return lo + h["offset"], hi + h["offset"]
    if k in ("rotate_x", "rotate_y", "rotate_z"):
        return h["bb"]
    if k == "medium":
        return bounding_box(h["boundary"])
    if k == "bvh":
        return h["min"], h["max"]
    raise ValueError(k)


def bvh_new(objects, rng):  # BVHNode::new src/accel.rs:98-136
    axis = rng.gen_range_u32(0, 3)  # drawn for every node, leaves included
    n = len(objects)
    if n == 1:
        bb = surrounding(bounding_box(objects[0]), bounding_box(objects[0]))
        return {"kind": "bvh", "min": bb[0], "max": bb[1], "left": objects[0], "right": objects[0], "single": True}
    if n == 2:
        a, b = bounding_box(objects[0]), bounding_box(objects[1])
        i1, i2 = (1, 0) if a[0][axis] < b[0][axis] else (0, 1)  # the larger min goes LEFT (:111-115)
        bb = surrounding(bounding_box(objects[i1]), bounding_box(objects[i2]))
        return {"kind": "bvh", "min": bb[0], "max": bb[1], "left": objects[i1], "right": objects[i2], "single": False}
    objects.sort(key=lambda o: float(bounding_box(o)[0][axis]))  # sort_by is stable, so is list.sort
    mid = n // 2
    left_part, right_part = objects[:mid], objects[mid:]
    left = bvh_new(left_part, rng)
    right = bvh_new(right_part, rng)
    objects[:mid], objects[mid:] = left_part, right_part  # the reference sorts its slice in place
    bb = surrounding((left["min"], left["max"]), (right["min"], right["max"]))
    return {"kind": "bvh", "min": bb[0], "max": bb[1], "left": left, "right": right, "single": False}


def camera_new(lookfrom, lookat, vup, vfov, aspect, aperture, focus_dist, time0, time1):  # src/main.rs:71-109
    theta = F(vfov) * F(math.pi / 180.0)
    h = F(np.tan(theta / F(2.0)))
    vh = h * F(2.0)
    vw = F(aspect) * vh
    w = unit(np.asarray(lookfrom, dtype=F) - np.asarray(lookat, dtype=F))
    u = unit(cross(np.asarray(vup, dtype=F), w))
    v = cross(w, u)
    origin = np.asarray(lookfrom, dtype=F)
    horizontal = u * vw * F(focus_dist)
    vertical = v * vh * F(focus_dist)
    llc = origin - horizontal / F(2.0) - vertical / F(2.0) - w * F(focus_dist)
    return {"origin": origin, "lower_left_corner": llc, "horizontal": horizontal, "vertical": vertical, "u": u, "v": v, "w": w,
            "lens_radius": F(aperture) / F(2.0), "time0": F(time0), "time1": F(time1)}


# ------------------------------------------------------------------------------------------------ the scenes
def _cornell_walls(white):  # src/scene.rs:645-672
    green, red = lambertian(solid(0.12, 0.45, 0.15)), lambertian(solid(0.65, 0.05, 0.05))
    return [flip_face(rect("yz", 0, 555, 0, 555, 555, green)), rect("yz", 0, 555, 0, 555, 0, red),
            flip_face(rect("xz", 0, 555, 0, 555, 0, white)), rect("xz", 0, 555, 0, 555, 555, white),
            flip_face(rect("xy", 0, 555, 0, 555, 555, white))]


def cornell_box(rng, assets_dir):  # src/scene.rs:630-730
    white = lambertian(solid(0.73, 0.73, 0.73))
    world = _cornell_walls(white)
    world.append(translate(rotate_y(boxy((0, 0, 0), (165, 330, 165), white), 15.0), (265, 0, 295)))
    world.append(sphere((190, 90, 190), 90.0, dielectric(1.5)))
    light_shape = rect("xz", 213, 343, 227, 332, 554, diffuse_light(solid(15, 15, 15)))
    world.append(flip_face(light_shape))
    cam = camera_new((278, 278, -800), (278, 278, 0), (0, 1, 0), 40.0, 1.0, 0.0, 10.0, 0.0, 1.0)
    return world, [light_shape], cam, 1.0


def cornell_smoke(rng, assets_dir):  # SURVEY 8(d) config 3, book 2's listing with the reference's constructors
    white = lambertian(solid(0.73, 0.73, 0.73))
    world = _cornell_walls(white)
    light_shape = rect("xz", 113, 443, 127, 432, 554, diffuse_light(solid(7, 7, 7)))
    world.append(flip_face(light_shape))
    world.append(constant_medium(translate(rotate_y(boxy((0, 0, 0), (165, 330, 165), white), 15.0), (265, 0, 295)), 0.01, solid(0, 0, 0)))
    world.append(constant_medium(translate(rotate_y(boxy((0, 0, 0), (165, 165, 165), white), -18.0), (130, 0, 65)), 0.01, solid(1, 1, 1)))
    cam = camera_new((278, 278, -800), (278, 278, 0), (0, 1, 0), 40.0, 1.0, 0.0, 10.0, 0.0, 1.0)
    return world, [light_shape], cam, 1.0


def random_spheres_demo(rng, assets_dir):  # src/scene.rs:167-284
    checker = {"kind": "checker", "odd": solid(0.1, 0.1, 0.1), "even": solid(0.9, 0.9, 0.9)}
    world = [sphere((0, -1000, 0), 1000.0, lambertian(checker))]
    for a in range(-11, 11):
        for b in range(-11, 11):
            choose_mat = rng.gen_f32()
            center = v3(F(a) + F(0.9) * rng.gen_f32(), F(0.2), F(b) + F(0.9) * rng.gen_f32())
            dd = center - v3(4.0, 0.2, 0.0)
            if np.sqrt(F(dd[0] * dd[0] + dd[1] * dd[1]) + dd[2] * dd[2]) > F(0.9):
                if choose_mat < F(0.8):
                    albedo = vrandom(rng) * vrandom(rng)
                    world.append(sphere(center, 0.2, lambertian(solid_v(albedo))))
                elif choose_mat < F(0.95):
                    albedo = vrandom_range(rng, 0.5, 1.0)
                    fuzz = rng.gen_range(0.0, 0.5)
                    world.append(sphere(center, 0.2, metal(solid_v(albedo), fuzz)))
                else:
                    world.append(sphere(center, 0.2, dielectric(1.5)))
    world.append(sphere((0, 1, 0), 1.0, dielectric(1.5)))
    world.append(sphere((-4, 1, 0), 1.0, lambertian(image_texture("assets/earthmap.png", assets_dir))))
    world.append(sphere((4, 1, 0), 1.0, metal(solid(0.7, 0.6, 0.5), 0.0)))
    light_shape = rect("xz", -11, 11, -11, 11, 8, diffuse_light(solid_v(v3(1.0, 0.77, 0.56) * F(2.0))))
    world.append(flip_face(light_shape))
    # RotatingCamera's first frame (:65-91, :254-281): angle 25, radius 20, height 2.5
    rad = F(25.0) * F(math.pi / 180.0)
    look = (F(20.0) * F(np.cos(rad)), F(2.5), F(20.0) * F(np.sin(rad)))
    cam = camera_new(look, (0, 1.5, 0), (0, 1, 0), 20.0, F(16.0) / F(9.0), 0.0, 10.0, 0.0, 1.0)
    return world, [light_shape], cam, F(16.0) / F(9.0)


def final_scene(rng, assets_dir):  # src/scene.rs:732-874
    ground = lambertian(solid(0.48, 0.83, 0.53))
    boxes1 = []
    for i in range(20):
        for j in range(20):
            w = F(100.0)
            x0, z0, y0 = F(-1000.0) + F(i) * w, F(-1000.0) + F(j) * w, F(0.0)
            x1, z1 = x0 + w, z0 + w
            y1 = rng.gen_range(1.0, 101.0)
            boxes1.append(boxy((x0, y0, z0), (x1, y1, z1), ground))
    objects = [bvh_new(boxes1, rng)]
    light_shape = rect("xz", 123, 423, 147, 412, 554, diffuse_light(solid(7, 7, 7)))
    objects.append(flip_face(light_shape))
    c1 = v3(400, 400, 200)
    objects.append(moving_sphere(c1, c1 + v3(30, 0, 0), 0.0, 1.0, 50.0, lambertian(solid(0.7, 0.3, 0.1))))
    objects.append(sphere((260, 150, 45), 50.0, dielectric(1.5)))
    objects.append(sphere((0, 150, 145), 50.0, metal(solid(0.8, 0.8, 0.9), 10.0)))
    boundary1 = sphere((360, 150, 145), 70.0, dielectric(1.5))
    objects.append(boundary1)
    objects.append(constant_medium(boundary1, 0.2, solid(0.2, 0.4, 0.9)))
    boundary2 = sphere((0, 0, 0), 5000.0, dielectric(1.5))
    objects.append(constant_medium(boundary2, 0.0001, solid(1, 1, 1)))
    objects.append(sphere((400, 200, 400), 100.0, lambertian(image_texture("assets/earthmap.png", assets_dir))))
    objects.append(sphere((220, 280, 300), 80.0, lambertian(noise_texture(rng, 0.1))))  # Perlin::new draws HERE
    white = lambertian(solid(0.73, 0.73, 0.73))
    boxes2 = [sphere(vrandom_range(rng, 0.0, 165.0), 10.0, white) for _ in range(1000)]
    objects.append(translate(rotate_y(bvh_new(boxes2, rng), 15.0), (-100, 270, 395)))
    cam = camera_new((478, 278, -600), (278, 278, 0), (0, 1, 0), 40.0, 1.0, 0.0, 10.0, 0.0, 1.0)
    return objects, [light_shape], cam, 1.0


def _demo_light_and_camera():  # the light and FixedCamera shared by balls_demo (:129-163) and perlin_demo (:303-336)
    light_shape = rect("xz", -6, 6, -6, 6, 8, diffuse_light(solid(4, 4, 4)))
    cam = camera_new((0, 2, 10), (0, 1, 0), (0, 1, 0), 40.0, F(16.0) / F(9.0), 0.0, 10.0, 0.0, 1.0)
    return light_shape, cam


def balls_demo(rng, assets_dir):  # src/scene.rs:93-165
    world = [sphere((0, 0, -1), 0.5, lambertian(solid(0.1, 0.2, 0.5))),
             sphere((0, -100.5, -1), 100.0, lambertian(solid(0.8, 0.8, 0.8))),
             sphere((1, 0, -1), 0.5, metal(solid(0.8, 0.6, 0.2), 0.3)),
             sphere((-1, 0, -1), 0.5, dielectric(1.5)),
             sphere((-1, 0, -1), -0.45, dielectric(1.5))]  # the hollow glass ball: negative radius
    light_shape, cam = _demo_light_and_camera()
    world.append(flip_face(light_shape))
    return world, [light_shape], cam, F(16.0) / F(9.0)


def perlin_demo(rng, assets_dir):  # src/scene.rs:286-337
    pertext = lambertian(noise_texture(rng, 2.0))  # Perlin::new draws here, before anything else
    world = [sphere((0, -1000, 0), 1000.0, pertext), sphere((0, 2, 0), 2.0, pertext)]
    light_shape, cam = _demo_light_and_camera()
    world.append(flip_face(light_shape))
    return world, [light_shape], cam, F(16.0) / F(9.0)


def _bowser(rng, assets_dir, x, y, z):  # Bowser::new src/scene.rs:344-541: 28 parts under a BVH of their own
    x, y, z = F(x), F(y), F(z)

    def X(d):  # `x - 2.0`, `x + 2.0`
        return x + F(d)

    def Y(d):  # `y - 1.875 + d`, evaluated left to right in f32
        return (y - F(1.875)) + F(d)

    def Z(d):  # `z + 4.5 - d`
        return (z + F(4.5)) - F(d)

    def img(name):
        return lambertian(image_texture("assets/" + name, assets_dir))

    parts = [rect("xy", X(-2.0), X(2.0), Y(1.0), Y(4.0), Z(3.0), img("bowser_face.png")),
             rect("xz", X(-2.0), X(2.0), Z(6.0), Z(3.0), Y(4.0), img("bowser_top.png")),
             rect("xy", X(-2.0), X(2.0), Y(1.0), Y(4.0), Z(6.0), img("bowser_back.png")),
             rect("yz", Y(1.0), Y(4.0), Z(6.0), Z(3.0), X(-2.0), img("bowser_side.png")),
             rect("yz", Y(1.0), Y(4.0), Z(6.0), Z(3.0), X(2.0), img("bowser_side.png"))]
    grey = lambertian(solid(0.278, 0.387, 0.438))
    parts.append(rect("xz", X(-2.0), X(2.0), Z(6.0), Z(3.0), Y(1.0), grey))
    brown = lambertian(solid(0.4, 0.2, 0.1))
    lightgrey = lambertian(solid(0.601, 0.687, 0.723))
    # (x0, y0, z0) - (x1, y1, z1) as the offsets the Rust adds to x, `y - 1.875` and subtracts from `z + 4.5`
    boxes = [
        # feet :424-443
        (grey, -1.5, 0.5, 4.75, -0.5, 1.0, 4.25), (grey, 0.5, 0.5, 4.75, 1.5, 1.0, 4.25),
        (grey, -1.5, 0.25, 4.75, -0.5, 0.5, 3.5), (grey, 0.5, 0.25, 4.75, 1.5, 0.5, 3.5),
        # arms :448-477
        (brown, -2.25, 1.75, 4.65, -2.00, 2.75, 4.35), (brown, -2.50, 1.75, 4.65, -2.25, 2.50, 4.35),
        (brown, -2.75, 1.75, 4.65, -2.50, 2.25, 4.35), (brown, 2.00, 1.75, 4.65, 2.25, 2.75, 4.35),
        (brown, 2.25, 1.75, 4.65, 2.50, 2.50, 4.35), (brown, 2.50, 1.75, 4.65, 2.75, 2.25, 4.35),
        # face rim :482-501
        (lightgrey, -2.0, 3.875, 3.00, 2.0, 4.00, 2.875), (lightgrey, -2.0, 1.0, 3.00, 2.0, 1.125, 2.875),
        (lightgrey, -2.0, 1.125, 3.00, -1.875, 3.875, 2.875), (lightgrey, 1.875, 1.125, 3.00, 2.0, 3.875, 2.875),
        # the eight port frames at the back :504-545
        (lightgrey, -1.875, 1.625, 6.125, -0.875, 1.75, 6.0), (lightgrey, -1.875, 1.125, 6.125, -0.875, 1.25, 6.0),
        (lightgrey, -1.875, 1.25, 6.125, -1.750, 1.625, 6.0), (lightgrey, -1.0, 1.25, 6.125, -0.875, 1.625, 6.0),
        (lightgrey, 0.875, 1.625, 6.125, 1.875, 1.75, 6.0), (lightgrey, 0.875, 1.125, 6.125, 1.875, 1.25, 6.0),
        (lightgrey, 1.750, 1.25, 6.125, 1.875, 1.625, 6.0), (lightgrey, 0.875, 1.25, 6.125, 1.0, 1.625, 6.0)]
    for m, x0, y0, z0, x1, y1, z1 in boxes:
        parts.append(boxy((X(x0), Y(y0), Z(z0)), (X(x1), Y(y1), Z(z1)), m))
    return bvh_new(parts, rng)  # `Bowser` itself only forwards hit / bounding_box to this node (:543-550)


def bowser_demo(rng, assets_dir):  # src/scene.rs:552-628
    checker = {"kind": "checker", "odd": solid(0.1, 0.1, 0.1), "even": solid(0.9, 0.9, 0.9)}
    world = [sphere((0, -1000, 0), 1000.0, lambertian(checker))]
    world.append(translate(rotate_x(rotate_y(rotate_z(_bowser(rng, assets_dir, 0.0, 0.0, 0.0), 0.0), 0.0), 0.0), (0.0, 1.625, -4.5)))
    light_shape = rect("xy", -2, 2, 1, 4, 3, diffuse_light(image_texture("assets/twitter.png", assets_dir)))
    world.append(flip_face(light_shape))
    # RotatingCamera's first frame (:65-91, :597-625): angle -35, radius 20, height 2.5
    rad = F(-35.0) * F(math.pi / 180.0)
    look = (F(20.0) * F(np.cos(rad)), F(2.5), F(20.0) * F(np.sin(rad)))
    cam = camera_new(look, (0, 2, 0), (0, 1, 0), 20.0, F(16.0) / F(9.0), 0.0, 10.0, 0.0, 1.0)
    return world, [light_shape], cam, F(16.0) / F(9.0)


SCENES = {"cornell_box": cornell_box, "cornell_smoke": cornell_smoke, "random_spheres_demo": random_spheres_demo, "final_scene": final_scene,
          "balls_demo": balls_demo, "perlin_demo": perlin_demo, "bowser_demo": bowser_demo}


# cam_iter of each builder: FixedCamera yields its camera once (src/scene.rs:24-46); RotatingCamera (:48-91) yields
# Camera::new from (radius cos a, height, radius sin a) while a <= limit, a += incr in f32
ROTATING = {"random_spheres_demo": dict(lookat=(0, 1.5, 0), vfov=20.0, height=2.5, angle=25.0, radius=20.0, incr=0.5, limit=360.0),   # :254-281
            "bowser_demo": dict(lookat=(0, 2, 0), vfov=20.0, height=2.5, angle=-35.0, radius=20.0, incr=0.5, limit=360.0 - 35.0)}  # :597-625


def cameras(name, seed=1, assets_dir=None):
    """Every camera `for cam in config.cam_iter` (src/main.rs:176) sees, in order."""
    if name not in ROTATING:
        yield SCENES[name](Rng(seed), assets_dir)[2]
        return
    r = ROTATING[name]
    angle, incr, limit = F(r["angle"]), F(r["incr"]), F(r["limit"])
    while not angle > limit:
        rad = angle * F(math.pi / 180.0)
        look = (F(r["radius"]) * F(np.cos(rad)), F(r["height"]), F(r["radius"]) * F(np.sin(rad)))
        yield camera_new(look, r["lookat"], (0, 1, 0), r["vfov"], F(16.0) / F(9.0), 0.0, 10.0, 0.0, 1.0)
        angle = angle + incr


def build(name, seed=1, assets_dir=None):
    """The reference's main(): scene builder, then BVHNode::new(&mut config.world) (src/main.rs:159-169)."""
    rng = Rng(seed)
    world, lights, cam, aspect = SCENES[name](rng, assets_dir)
    return {"world": bvh_new(world, rng), "lights": lights, "camera": cam, "aspect_ratio": F(aspect)}


# ------------------------------------------------------------------------------------------------ the product's scene, read back
def unlower(desc, cam=None, aspect=None):
    """vk_scene_desc (include/vecchio_gpu.h) -> the same nested form.  Index arithmetic only; no product code runs."""
    def f3(a):
        return np.array([a[0], a[1], a[2]], dtype=F)

    def tex(i):
        t = desc.textures[i]
        w = [int(t.w[0]), int(t.w[1]), int(t.w[2])]
        if t.type == 0:
            return {"kind": "solid", "rgb": np.array(w, dtype=np.uint32).view(F)}
        if t.type == 1:
            return {"kind": "checker", "odd": tex(w[0]), "even": tex(w[1])}
        if t.type == 2:
            off, wd, ht = w
            texels = np.ctypeslib.as_array(desc.texels, shape=(int(desc.n_texel_bytes),))[off:off + wd * ht * 3]
            return {"kind": "image", "width": wd, "height": ht, "texels": np.array(texels, dtype=np.uint8)}
        p = desc.perlins[w[0]]
        return {"kind": "noise", "scale": np.uint32(w[1]).view(F),
                "perlin": {"ranvec": np.array(p.ranvec, dtype=F).reshape(256, 3), "perm_x": np.array(p.perm_x, dtype=np.uint8),
                           "perm_y": np.array(p.perm_y, dtype=np.uint8), "perm_z": np.array(p.perm_z, dtype=np.uint8)}}

    def mat(i):
        m = desc.materials[i]
        if m.type == 0:
            return lambertian(tex(m.tex))
        if m.type == 1:
            return metal(tex(m.tex), m.param)
        if m.type == 2:
            return dielectric(m.param)
        if m.type == 3:
            return diffuse_light(tex(m.tex))
        if m.type == 4:
            return isotropic(tex(m.tex))
        return {"kind": "specdiffuse", "specular": mat(m.tex), "diffuse": mat(m.aux), "pct": F(m.param)}

    def hit(ref):
        t, i = ref >> 28, ref & 0x0FFFFFFF
        if t == 1:
            n = desc.nodes[i]
            left = hit(n.left)
            single = n.left == n.right
            return {"kind": "bvh", "min": f3(n.bb_min), "max": f3(n.bb_max), "left": left, "right": left if single else hit(n.right), "single": single}
        if t == 2:
            s = desc.spheres[i]
            return sphere(f3(s.center), s.radius, mat(desc.sphere_mat[i]))
        if t == 3:
            s = desc.mspheres[i]
            return moving_sphere(f3(s.center0), f3(s.center1), s.time0, s.time1, s.radius, mat(s.mat))
        if t == 4:
            r = desc.rects[i]
            return {"kind": "rect", "c0": F(r.c0), "c1": F(r.c1), "d0": F(r.d0), "d1": F(r.d1), "k": F(r.k),
                    "axes": (r.axes & 3, (r.axes >> 2) & 3, (r.axes >> 4) & 3), "flip": bool(r.axes & 0x100), "mat": mat(r.mat)}
        if t == 5:
            b = desc.boxes[i]
            return {"kind": "box", "min": f3(b.box_min), "max": f3(b.box_max), "mat": mat(b.mat)}
        if t == 6:
            x = desc.xforms[i]
            if x.kind == 0:
                return {"kind": "translate", "offset": v3(x.a, x.b, x.c), "child": hit(x.child)}
            if x.kind == 4:
                return {"kind": "flip", "child": hit(x.child)}
            return {"kind": {1: "rotate_x", 2: "rotate_y", 3: "rotate_z"}[x.kind], "sin": F(x.a), "cos": F(x.b), "child": hit(x.child)}
        if t == 7:
            m = desc.media[i]
            return {"kind": "medium", "boundary": hit(m.boundary), "neg_inv_density": F(m.neg_inv_density), "mat": mat(m.mat)}
        raise ValueError(f"bad ref type {t}")

    out = {"world": hit(desc.root), "lights": [hit(desc.lights[i]) for i in range(desc.n_lights)]}
    if cam is not None:
        out["camera"] = {k: (np.array(getattr(cam, k)[:], dtype=F) if k not in ("lens_radius", "time0", "time1") else F(getattr(cam, k)))
                         for k in ("origin", "lower_left_corner", "horizontal", "vertical", "u", "v", "w", "lens_radius", "time0", "time1")}
    if aspect is not None:
        out["aspect_ratio"] = F(aspect)
    return out


# keys whose value passes through libm (sin / cos / tan) before it is stored: numpy's f32 kernels and glibc's may differ in the
# last bit, everything else is +, -, *, /, sqrt and must agree exactly
LIBM_KEYS = {"sin", "cos", "origin", "lower_left_corner", "horizontal", "vertical", "u", "v", "w", "bb"}


def compare(a, b, path="scene", libm=False, stats=None):
    """Raises AssertionError naming the first differing path; returns the number of leaves compared."""
    stats = stats if stats is not None else {"leaves": 0, "exact": 0}
    if isinstance(a, dict):
        assert isinstance(b, dict) and a.get("kind") == b.get("kind"), f"{path}: kind {a.get('kind')} != {b.get('kind') if isinstance(b, dict) else b}"
        for k in a:
            if k == "bb":  # the front end's cached rotate box: checked through the BVH node boxes above it
                continue
            assert k in b, f"{path}.{k} missing"
            # a BVH box above a rotation inherits the rotation's last-bit freedom
            loose = libm or k in LIBM_KEYS or (k in ("min", "max") and a.get("kind") == "bvh")
            compare(a[k], b[k], f"{path}.{a.get('kind', '')}.{k}" if "kind" in a else f"{path}.{k}", loose, stats)
        return stats
    if isinstance(a, (list, tuple)) and not isinstance(a, np.ndarray):
        assert len(a) == len(b), f"{path}: length {len(a)} != {len(b)}"
        for i, (x, y) in enumerate(zip(a, b)):
            compare(x, y, f"{path}[{i}]", libm, stats)
        return stats
    x, y = np.asarray(a), np.asarray(b)
    assert x.shape == y.shape, f"{path}: shape {x.shape} != {y.shape}"
    stats["leaves"] += x.size
    if x.dtype.kind == "f" or y.dtype.kind == "f":
        x, y = x.astype(F), y.astype(F)
        if np.array_equal(x, y):
            stats["exact"] += x.size
        else:
            assert libm, f"{path}: {x.tolist()} != {y.tolist()} (must be bit-identical)"
            assert np.allclose(x, y, rtol=4e-7, atol=4e-7 * float(np.abs(x).max())), f"{path}: {x.tolist()} !~ {y.tolist()}"
    else:
        assert np.array_equal(x, y), f"{path}: {x.reshape(-1)[:8].tolist()} != {y.reshape(-1)[:8].tolist()}"
        stats["exact"] += x.size
    return stats
