"""The warp-queue and step-queue kernels' own protocol check (compute-sanitizer is closed on this pool: its racecheck / memcheck cannot
run).  A debug build of vk_warpq.cu (-DVKQ_SELFCHECK=1) verifies before EVERY scheduling decision that every queued
index is a valid slot, that no slot is queued twice, and that queues plus rays in flight never exceed the pool; the
violation count must be zero on flat, media and BVH scenes, the pools must drain (the render returns), and the frames
must equal the normal build's bit for bit."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "build", "libvk_selfcheck.so")

CHILD = r"""
import ctypes as C, hashlib, sys
sys.path.insert(0, %r)
import numpy as np
import vecchio_b200 as vb
L = vb.gpu_lib()
L.vk_debug_counters.argtypes = [C.c_void_p, C.POINTER(C.c_ulonglong)]
ctx = vb.Context(0)
WQ, SQ = vb.VK_VARIANT_WARPQ, vb.VK_VARIANT_STEPQ
for name, W, spp, depth, flags, variant in (("cornell_box", 97, 24, 100, 0, WQ), ("cornell_smoke", 64, 16, 100, 0, WQ), ("final_scene", 48, 8, 50, 0, WQ),
                                            ("random_spheres_demo", 80, 8, 50, 0, WQ), ("cornell_box", 64, 8, 100, vb.VK_FLAG_LEGACY_SCATTER, WQ),
                                            ("cornell_box", 33, 7, 100, vb.VK_FLAG_STRICT_MATH, WQ),
                                            ("final_scene", 48, 8, 50, 0, SQ), ("random_spheres_demo", 80, 8, 50, 0, SQ),
                                            ("stress_spheres", 64, 4, 50, 0, SQ), ("final_scene", 40, 4, 50, vb.VK_FLAG_LEGACY_SCATTER, SQ)):
    s = vb.Scene(name, param=60 if name == "stress_spheres" else 0); cam = s.next_camera(); ctx.upload(s)
    rgb, _, st = ctx.render(cam, vb.render_params(W, s.height_for(W), spp, depth, seed=3, variant=variant, flags=flags))
    out = (C.c_ulonglong * 8)(); L.vk_debug_counters(ctx._h, out)
    print(name + ":" + str(variant), flags, st.paths, st.rays, int(out[5]), hashlib.sha256(rgb.tobytes()).hexdigest()[:16], flush=True)
"""


def run_child(lib):
    env = dict(os.environ)
    if lib:
        env["VECCHIO_GPU_LIB"] = lib
    r = subprocess.run([sys.executable, "-c", CHILD % ROOT], capture_output=True, text=True, env=env, cwd=ROOT, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    return [line.split() for line in r.stdout.strip().splitlines()]


@pytest.mark.gpu
def test_queue_protocol_selfcheck_finds_nothing_and_changes_nothing():
    if not os.path.exists(LIB):
        subprocess.run([os.path.join(ROOT, "scripts", "build_variants.sh"), "selfcheck:-DVKQ_SELFCHECK=1"], check=True, cwd=ROOT,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    checked, normal = run_child(LIB), run_child(None)
    assert len(checked) == len(normal) == 10
    for c, n in zip(checked, normal):
        assert int(c[4]) == 0, ("queue protocol violations", c)
        assert c[:4] == n[:4] and c[5] == n[5], ("the self-checking build renders another frame", c, n)
    # the render build of the strict-math frame is only compiled without the check: its line must simply agree


@pytest.mark.gpu
def test_step_queue_overflow_stack_renders_the_same_frames():
    """The step-queue kernel keeps the first VKS_SD (4) stack entries of a traversal in shared memory and the rest in a
    per-warp strip of global memory.  A build with VKS_SD=1 pushes almost every stack access of these scenes through
    the overflow path: same rays, same frames."""
    lib = os.path.join(ROOT, "build", "libvk_sd1.so")
    if not os.path.exists(lib):
        subprocess.run([os.path.join(ROOT, "scripts", "build_variants.sh"), "sd1:-DVKS_SD=1"], check=True, cwd=ROOT,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    shallow, normal = run_child(lib), run_child(None)
    assert len(shallow) == len(normal) == 10
    for c, n in zip(shallow, normal):
        assert c[:4] == n[:4] and c[5] == n[5], ("the overflow stack changes the frame", c, n)
