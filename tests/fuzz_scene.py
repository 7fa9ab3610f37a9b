"""Corrupts lowered scene descriptions at random and hands them to vk_scene_check (the validator and layout
planner vk_scene_upload runs before it touches the device).  Run as a subprocess by test_scene_fuzz.py so that a
crash in the native library fails one test instead of taking pytest down.

    python tests/fuzz_scene.py SCENE SEED COUNT

Array COUNTS are only ever reduced: a count larger than the caller's array is the caller's out-of-bounds read and
cannot be detected by any callee."""
import ctypes as C
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vecchio_b200 as vb

ARRAYS = [("nodes","n_nodes"),("spheres","n_spheres"),("sphere_mat","n_spheres"),("mspheres","n_mspheres"),("rects","n_rects"),("boxes","n_boxes"),
          ("xforms","n_xforms"),("media","n_media"),("lights","n_lights"),("materials","n_materials"),("textures","n_textures")]
def nasty_u32(rng, d):
    k = rng.randrange(6)
    if k == 0: return 0
    if k == 1: return 0xFFFFFFFF
    if k == 2: return (rng.randrange(16) << 28) | rng.randrange(1 << 28)
    if k == 3: return (rng.randrange(1, 8) << 28) | rng.randrange(0, 8)
    if k == 4: return rng.randrange(0, 40)
    return rng.randrange(1 << 32)
def nasty_f32(rng):
    return rng.choice([0.0, -0.0, float('nan'), float('inf'), -float('inf'), 1e38, -1e38, 1e-45, rng.uniform(-1e3, 1e3)])
def set_leaf(rng, obj, d):
    fields = obj._fields_
    name, typ = rng.choice(fields)[:2]
    cur = getattr(obj, name)
    if isinstance(cur, C.Array):
        i = rng.randrange(len(cur))
        if isinstance(cur[i], float): cur[i] = nasty_f32(rng)
        elif isinstance(cur[i], int): cur[i] = nasty_u32(rng, d) & ((1 << (8*C.sizeof(cur._type_))) - 1)
        elif isinstance(cur[i], C.Array):
            j = rng.randrange(len(cur[i])); cur[i][j] = nasty_f32(rng) if isinstance(cur[i][j], float) else 0
        return f"{name}[{i}]"
    if isinstance(cur, float): setattr(obj, name, nasty_f32(rng))
    elif isinstance(cur, int): setattr(obj, name, nasty_u32(rng, d) & ((1 << (8*C.sizeof(typ))) - 1))
    elif isinstance(cur, (C.Structure, C.Union)): return name + "." + set_leaf(rng, cur, d)
    return name
def mutate(rng, d):
    what = rng.randrange(10)
    if what == 0:
        d.root = nasty_u32(rng, d); return "root"
    if what == 1:
        arr, cnt = rng.choice(ARRAYS); n = getattr(d, cnt)
        if n: setattr(d, cnt, rng.randrange(n)); return f"{cnt} {n}->{getattr(d,cnt)}"
        return "noop"
    arr, cnt = rng.choice(ARRAYS); n = getattr(d, cnt)
    if n == 0: return "noop"
    i = rng.randrange(n); a = getattr(d, arr)
    if isinstance(a[i], int):
        a[i] = nasty_u32(rng, d); return f"{arr}[{i}]"
    return f"{arr}[{i}]." + set_leaf(rng, a[i], d)

def main(scene_name, seed0, n):
    stats = {}
    for k in range(n):
        rng = random.Random(seed0 * 100003 + k)
        s = vb.Scene(scene_name, param=12 if scene_name == 'stress_spheres' else 0)
        d = s.desc
        desc = [mutate(rng, d) for _ in range(rng.choice([1, 1, 1, 2, 3]))]
        print(scene_name, k, desc, flush=True)  # the last line before a crash names the culprit
        try:
            info = vb.scene_check(s.desc_ptr); rc = 0
        except vb.VecchioError as e:
            rc = e.code
        stats[rc] = stats.get(rc, 0) + 1
    print("DONE", scene_name, stats, flush=True)
if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]))
