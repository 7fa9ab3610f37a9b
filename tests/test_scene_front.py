"""The product's scene path (vecchio_b200/host/scene.cpp builders -> BVHNode::new -> every lower()) against the oracle's
own front end (oracle/scene_front.py: the reference's scene.rs / accel.rs / hittable.rs bounding boxes / Perlin::new /
Camera::new restated independently in Python).  Record by record: every constant, every seeded random draw in the
reference's draw order, every BVH split and box, the decoded texture bytes (a different PNG decoder), the camera.
Without this, a wrong constant in scene.cpp or a wrong field in a lower() would be common to both sides of every
oracle-vs-GPU comparison (the C++ oracle rebuilds its objects from the flattened scene)."""
import copy

import numpy as np
import pytest

from conftest import get_scene
from oracle import scene_front as sf

NAMES = ["cornell_box", "cornell_smoke", "random_spheres_demo", "final_scene", "balls_demo", "perlin_demo", "bowser_demo"]


@pytest.fixture(scope="module")
def pairs(vb):
    out = {}
    for name in NAMES:
        scene, cam = get_scene(vb, name, seed=1)
        out[name] = (sf.build(name, seed=1, assets_dir=vb.ASSETS_DIR), sf.unlower(scene.desc, cam, scene.aspect_ratio), scene)
    return out


@pytest.mark.parametrize("name", NAMES)
def test_lowered_scene_equals_the_independent_front_end(pairs, name):
    front, product, scene = pairs[name]
    st = sf.compare(front, product)
    # almost everything is bit-identical; only values that went through sin / cos / tan may differ in the last place
    assert st["leaves"] > 60 and st["exact"] >= 0.98 * st["leaves"], st
    c = scene.census()
    if name == "cornell_box":
        assert (c["nodes"], c["rects"], c["boxes"], c["spheres"], c["xforms"], c["lights"]) == (7, 7, 1, 1, 2, 1)
    if name == "final_scene":
        assert (c["nodes"], c["boxes"], c["spheres"], c["mspheres"], c["media"], c["perlins"]) == (13 + 511 + 1023, 400, 1006, 1, 2, 1)


def test_another_seed_is_another_scene_and_still_agrees(vb):
    scene = vb.Scene("random_spheres_demo", seed=7)
    cam = scene.next_camera()
    sf.compare(sf.build("random_spheres_demo", seed=7, assets_dir=vb.ASSETS_DIR), sf.unlower(scene.desc, cam, scene.aspect_ratio))
    with pytest.raises(AssertionError):
        sf.compare(sf.build("random_spheres_demo", seed=1, assets_dir=vb.ASSETS_DIR), sf.unlower(scene.desc, cam, scene.aspect_ratio))


def _first(tree, kind):
    stack = [tree]
    while stack:
        n = stack.pop()
        if isinstance(n, dict):
            if n.get("kind") == kind:
                return n
            stack.extend(n.values())
        elif isinstance(n, list):
            stack.extend(n)
    raise KeyError(kind)


@pytest.mark.parametrize("kind,key", [("rect", "k"), ("rect", "c1"), ("box", "max"), ("sphere", "radius"), ("translate", "offset"),
                                      ("rotate_y", "sin"), ("dielectric", "ior"), ("solid", "rgb"), ("bvh", "min")])
def test_the_comparison_notices_one_changed_value(pairs, kind, key):
    """What the check is for: one constant of scene.cpp, or one field a lower() writes, off by a little."""
    front, product, _ = pairs["cornell_box"]
    bad = copy.deepcopy(product)
    node = _first(bad["world"], kind)
    node[key] = (np.asarray(node[key], dtype=np.float32) * np.float32(1.0001) + np.float32(1e-4)).astype(np.float32)
    with pytest.raises(AssertionError):
        sf.compare(front, bad)


def test_the_comparison_notices_structure_changes(pairs):
    front, product, _ = pairs["cornell_box"]
    bad = copy.deepcopy(product)
    r = _first(bad["world"], "rect")
    r["flip"] = not r["flip"]
    with pytest.raises(AssertionError):
        sf.compare(front, bad)
    bad = copy.deepcopy(product)
    n = _first(bad["world"], "bvh")
    n["left"], n["right"] = n["right"], n["left"]
    with pytest.raises(AssertionError):
        sf.compare(front, bad)
    bad = copy.deepcopy(pairs["final_scene"][1])
    p = _first(bad["world"], "noise")["perlin"]
    p["perm_y"][3], p["perm_y"][4] = p["perm_y"][4], p["perm_y"][3]
    with pytest.raises(AssertionError):
        sf.compare(pairs["final_scene"][0], bad)


@pytest.mark.parametrize("name,frames", [("random_spheres_demo", 671), ("bowser_demo", 721), ("cornell_box", 1), ("perlin_demo", 1)])
def test_camera_iterator_yields_the_reference_s_frames(vb, name, frames):
    """`for cam in config.cam_iter` (src/main.rs:176): FixedCamera once, RotatingCamera 671 / 721 turntable frames
    (src/scene.rs:24-91, 254-281, 597-625) -- the product's iterator against the front end's, frame by frame."""
    scene = vb.Scene(name, seed=1)
    keys = ("origin", "lower_left_corner", "horizontal", "vertical", "u", "v", "w")
    n = 0
    for want in sf.cameras(name, seed=1, assets_dir=vb.ASSETS_DIR):
        cam = scene.next_camera()
        assert cam is not None, f"{name}: the product's iterator ended after {n} frames"
        for k in keys:
            got = np.array(getattr(cam, k)[:], dtype=np.float32)
            assert np.allclose(got, want[k], rtol=2e-6, atol=2e-6 * float(np.abs(want[k]).max()) + 1e-6), f"{name} frame {n}: {k} {got} != {want[k]}"
        assert (cam.lens_radius, cam.time0, cam.time1) == (float(want["lens_radius"]), float(want["time0"]), float(want["time1"]))
        n += 1
    assert n == frames and scene.next_camera() is None
