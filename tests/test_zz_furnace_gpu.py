"""The furnace closed forms (tests/test_oracle.py) on the GPU path.

Round 1 shipped these behind a non-strict xfail and the driver's B200 run XFAILED the white-medium one.  Measured
since (scripts/diag_furnace.py, profiles/r2_diag_furnace.log): on every variant and both math builds the medium
furnace drops no sample, its mean is E within 1.1 sigma (ratio 1.0019 / 1.0011 on two seeds) and the legacy
integrator is exact to 4e-8 -- the assert that tripped was the sanity bound `rays > 1.2 * paths` copied from the
oracle test: the oracle counts the segments the reference keeps tracing after a bounce weight of exactly 0
(`rays`), the GPU ends such a path (the comparable oracle figure is `rays_live`) and traces 1.182
segments per path here.  The bound is now 1.1 and the marker is gone."""
import numpy as np
import pytest

from conftest import get_scene
from test_oracle import FURNACE_E

pytestmark = [pytest.mark.gpu]


def test_furnace_lambertian_closed_form_gpu(vb, ctx):
    scene, cam = get_scene(vb, "furnace_demo", param=0)
    ctx.upload(scene)
    W, spp = 128, 1024
    rgb, sq, st = ctx.render(cam, vb.render_params(W, W, spp, 100, seed=7), want_sumsq=True)
    assert st.dropped_samples == 0
    rgb = rgb.astype(np.float64)
    assert np.allclose(rgb[:8, :8], FURNACE_E, rtol=1e-5)
    yy, xx = np.mgrid[0:W, 0:W]
    disc = ((yy - 63.5) ** 2 + (xx - 63.5) ** 2) < 24 ** 2
    want = np.array([0.5, 0.25, 0.75]) * FURNACE_E
    mean = rgb[disc].mean(axis=0)
    sigma = np.sqrt((sq[disc] / spp - rgb[disc] ** 2).mean(axis=0) / spp / disc.sum())
    assert np.all(np.abs(mean - want) <= 4 * sigma) and np.all(np.abs(mean / want - 1) < 2e-3), (mean / want, (mean - want) / sigma)


@pytest.mark.parametrize("kind,albedo", [(1, (0.7, 0.6, 0.5)), (2, (1.0, 1.0, 1.0))])
def test_furnace_specular_closed_form_gpu(vb, ctx, kind, albedo):
    scene, cam = get_scene(vb, "furnace_demo", param=kind)
    ctx.upload(scene)
    W = 96
    rgb, _, st = ctx.render(cam, vb.render_params(W, W, 64, 100, seed=3))
    assert st.dropped_samples == 0
    want = np.array(albedo) * FURNACE_E
    assert np.allclose(rgb[40:56, 40:56], want, rtol=5e-5)
    assert np.allclose(rgb[:8, :8], FURNACE_E, rtol=1e-5)


def test_furnace_white_medium_conserves_energy_gpu(vb, ctx):
    scene, cam = get_scene(vb, "furnace_demo", param=3)
    ctx.upload(scene)
    W, spp = 128, 2048
    rgb, sq, st = ctx.render(cam, vb.render_params(W, W, spp, 100, seed=1), want_sumsq=True)
    assert st.dropped_samples == 0
    assert st.rays > 1.1 * st.paths  # the medium does scatter (1.18 segments per path; zero-weight continuations are not traced)
    rgb = rgb.astype(np.float64)
    yy, xx = np.mgrid[0:W, 0:W]
    disc = ((yy - 63.5) ** 2 + (xx - 63.5) ** 2) < 24 ** 2
    mean = rgb[disc].mean(axis=0)
    sigma = np.sqrt((sq[disc] / spp - rgb[disc] ** 2).mean(axis=0) / spp / disc.sum())
    assert np.all(np.abs(mean - FURNACE_E) <= 4 * sigma) and np.all(np.abs(mean / FURNACE_E - 1) < 0.01), (mean / FURNACE_E, sigma)
    legacy, _, _ = ctx.render(cam, vb.render_params(W, W, 16, 100, seed=1, flags=vb.VK_FLAG_LEGACY_SCATTER))
    assert np.allclose(legacy, FURNACE_E, rtol=1e-5)
