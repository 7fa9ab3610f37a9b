"""The callers either side of the path on the GPU: the frame-loop program
(vecchio_b200/host/main.cpp, the reference's main() over the C ABI) and the additive checkpoint of sample
sums, both against the calls the parity tests already cover."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, get_scene
from test_frames_and_checkpoints import RENDER_BIN, parse_ppm

pytestmark = pytest.mark.gpu


def run_program(tmp_path, *args):
    out = tmp_path / "frames"
    out.mkdir(exist_ok=True)
    r = subprocess.run([RENDER_BIN, "--out-dir", str(out)] + [str(a) for a in args], capture_output=True, text=True,
                       cwd=ROOT, timeout=600)
    assert r.returncode == 0, r.stderr
    return out, r.stderr


def test_program_writes_the_frame_render_rgb8_returns(vb, ctx, tmp_path):
    # main.rs defaults: scene 1 = cornell_box, one camera, file output_0000.ppm
    out, err = run_program(tmp_path, "--width", 96, "--spp", 32, "--depth", 50, "--seed", 3)
    assert "Generating scene..." in err and "Wrote frame" in err
    assert sorted(os.listdir(out)) == ["output_0000.ppm"]  # cornell_box's iterator holds one camera
    scene, cam = get_scene(vb, "cornell_box")
    ctx.upload(scene)
    rgb8, st = ctx.render_rgb8(cam, vb.render_params(96, scene.height_for(96), 32, 50, seed=3))
    assert (parse_ppm(str(out / "output_0000.ppm")) == rgb8).all()  # same samples, same conversion: bit for bit
    assert rgb8.mean() > 20  # a lit room, not a black frame


def test_program_turntable_keeps_the_scene_resident(vb, ctx, tmp_path):
    # scene 3 = random_spheres_demo, RotatingCamera: one file per camera, numbering follows the iterator
    out, err = run_program(tmp_path, "--scene", 3, "--width", 80, "--spp", 8, "--depth", 20, "--first-frame", 1,
                           "--frames", 2)
    assert sorted(os.listdir(out)) == ["output_0001.ppm", "output_0002.ppm"] and err.count("Wrote frame") == 2
    scene = vb.Scene("random_spheres_demo", seed=1)
    scene.next_camera()
    cam1 = scene.next_camera()
    ctx.upload(scene)
    rgb8, _ = ctx.render_rgb8(cam1, vb.render_params(80, scene.height_for(80), 8, 20, seed=1 + 1))
    f1, f2 = parse_ppm(str(out / "output_0001.ppm")), parse_ppm(str(out / "output_0002.ppm"))
    assert (f1 == rgb8).all()
    assert (f1 != f2).mean() > 0.05  # the camera moved


def test_program_legacy_sky_scene(vb, ctx, tmp_path):
    out, _ = run_program(tmp_path, "--scene", "random_spheres_cover", "--legacy", "--sky", "--width", 64, "--spp", 8,
                         "--depth", 20)
    scene, cam = get_scene(vb, "random_spheres_cover")
    ctx.upload(scene)
    flags = vb.VK_FLAG_LEGACY_SCATTER | vb.VK_FLAG_SKY_BACKGROUND
    rgb8, _ = ctx.render_rgb8(cam, vb.render_params(64, scene.height_for(64), 8, 20, seed=1, flags=flags))
    assert (parse_ppm(str(out / "output_0000.ppm")) == rgb8).all()


def test_program_refuses_what_the_library_refuses(tmp_path):
    # a lightless scene with the HEAD integrator: the reference panics at src/hittable.rs:431; no file is written
    out = tmp_path / "frames"
    out.mkdir()
    r = subprocess.run([RENDER_BIN, "--scene", "random_spheres_cover", "--width", "64", "--spp", "4", "--out-dir", str(out)],
                       capture_output=True, text=True, cwd=ROOT, timeout=600)
    assert r.returncode == 1 and "light" in r.stderr and not os.listdir(out)


def test_program_on_two_gpus_in_one_process(vb, ctx, tmp_path):
    try:
        vb.Context(1).close()
    except vb.VecchioError:
        pytest.skip("one GPU")
    out1, _ = run_program(tmp_path, "--width", 96, "--spp", 32, "--depth", 50)
    f1 = parse_ppm(str(out1 / "output_0000.ppm"))
    out2, _ = run_program(tmp_path, "--width", 96, "--spp", 32, "--depth", 50, "--gpus", 2)
    f2 = parse_ppm(str(out2 / "output_0000.ppm"))
    # the same Philox samples into integer accumulators, the peers' added on GPU 0 over NVLink: the same frame, bit for bit
    assert (f1 == f2).all()


def test_multi_context_frame_equals_the_one_gpu_frame(vb, ctx):
    """vk_multi_render (spp slices per device, peer-memory reduce on device 0) against vk_render on one device:
    identical floats, identical segment counts, for a range that does not divide evenly and with more devices than
    samples in the slice."""
    try:
        m = vb.MultiContext([0, 1])
    except vb.VecchioError:
        pytest.skip("one GPU")
    scene, cam = get_scene(vb, "cornell_smoke")
    ctx.upload(scene)
    m.upload(scene)
    for kw in (dict(spp=33), dict(spp=40, spp_begin=7, spp_count=21), dict(spp=8, spp_begin=3, spp_count=1)):
        p = vb.render_params(96, 96, max_depth=100, seed=4, **kw)
        a, qa, sa = ctx.render(cam, p, want_sumsq=True)
        b, qb, sb = m.render(cam, p, want_sumsq=True)
        assert np.array_equal(a, b) and np.allclose(qa, qb, rtol=1e-6) and (sa.paths, sa.rays) == (sb.paths, sb.rays)
    r1, _ = ctx.render_rgb8(cam, vb.render_params(96, 96, 33, 100, seed=4))
    r2, _ = m.render_rgb8(cam, vb.render_params(96, 96, 33, 100, seed=4))
    assert np.array_equal(r1, r2)
    m.close()


def test_slices_merge_into_the_frame(vb, ctx, tmp_path):
    from vecchio_b200 import accumulate
    scene, cam = get_scene(vb, "cornell_box")
    ctx.upload(scene)
    W, SPP = 128, 64
    p = vb.render_params(W, W, SPP, 100, seed=9)
    full, full_sq, st = ctx.render(cam, p, want_sumsq=True)
    key = accumulate.frame_key("cornell_box", 1, cam, p)
    path = str(tmp_path / "frame.npz")
    acc = accumulate.FrameAccumulator(key, with_sumsq=True)
    assert accumulate.render_resumable(ctx, cam, p, acc, slice_spp=24, checkpoint_path=path, max_slices=1) == 1
    acc = accumulate.FrameAccumulator.load(path)  # a second run picks the frame up where the first stopped
    assert acc.missing() == [(24, 64)]
    assert accumulate.render_resumable(ctx, cam, p, acc, slice_spp=24, checkpoint_path=path) == 2
    assert acc.complete and acc.paths == st.paths and acc.rays == st.rays
    assert np.allclose(acc.frame(), full, rtol=1e-5, atol=1e-6)
    assert np.allclose(acc.sumsq, full_sq, rtol=1e-5, atol=1e-6)
    se = acc.standard_error()
    assert np.isfinite(se).all() and 0 < se.mean() < 0.2


def test_plain_c_example_renders_the_same_frame(vb, ctx, tmp_path):
    """examples/minimal.c (the ABI from C11) against the same calls made through ctypes."""
    from test_host_and_abi import build_minimal_c
    exe = build_minimal_c(tmp_path)
    out = tmp_path / "c.ppm"
    r = subprocess.run([exe, str(out)], capture_output=True, text=True, cwd=ROOT, timeout=300)
    assert r.returncode == 0, r.stderr
    scene, cam = get_scene(vb, "cornell_box")
    ctx.upload(scene)
    rgb8, _ = ctx.render_rgb8(cam, vb.render_params(200, 200, 64, 100, seed=1))
    assert (parse_ppm(str(out)) == rgb8).all()
