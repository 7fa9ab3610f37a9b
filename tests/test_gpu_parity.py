"""GPU parity tests (run on a B200 with `pytest -m gpu`).  Everything goes through the C ABI
(libvecchio_gpu.so); the oracle is only the checker.

Bars (BASELINE.json north_star):
  * hit parity: closest-hit primitive ids bit-exact (ties excepted), t and normal within 1e-5
    relative, on identical ray batches (camera rays + secondary rays harvested from oracle paths);
  * image parity: per pixel and channel within 3 sigma of the Monte-Carlo standard error.
"""
import ctypes as C
import json
import os

import numpy as np
import pytest

from conftest import ROOT, get_scene

pytestmark = pytest.mark.gpu
INF = np.float32(np.inf)

# (scene, param, image width used for ray generation, rays per batch)
HIT_SCENES = [("cornell_box", 0, 200, 200_000), ("cornell_smoke", 0, 200, 200_000), ("random_spheres_demo", 0, 200, 200_000),
              ("final_scene", 0, 200, 200_000), ("bowser_demo", 0, 200, 100_000), ("perlin_demo", 0, 160, 50_000),
              ("balls_demo", 0, 160, 50_000), ("stress_spheres", 64, 200, 100_000), ("api_surface_demo", 0, 160, 50_000)]


def camera_rays(cam, n, rng):
    """Camera::get_ray (src/main.rs:111-120) for aperture-0 cameras, vectorised in fp32."""
    s = rng.random(n, dtype=np.float32)
    t = rng.random(n, dtype=np.float32)
    f = lambda v: np.array(list(v), dtype=np.float32)  # noqa: E731
    rays = np.zeros(n, dtype=[("origin", "<f4", 3), ("direction", "<f4", 3), ("time", "<f4"), ("tmin", "<f4"), ("tmax", "<f4")])
    rays["origin"] = f(cam.origin)
    rays["direction"] = (f(cam.lower_left_corner)[None] + s[:, None] * f(cam.horizontal)[None] + t[:, None] * f(cam.vertical)[None]
                         - f(cam.origin)[None]).astype(np.float32)
    rays["time"] = rng.random(n, dtype=np.float32)
    rays["tmin"], rays["tmax"] = 0.001, INF
    return rays


def ray_batch(vb, o, scene, cam, width, n, seed):
    rng = np.random.default_rng(seed)
    a = camera_rays(cam, n // 2, rng)
    b = o.harvest_rays(cam, width, scene.height_for(width), 50, seed, n - len(a))
    return np.concatenate([a, b.astype(a.dtype)])


def compare_hits(vb, ref, got, t_rtol, n_tol, uv_tol, label):
    hit_r, hit_g = ref["prim"] != 0, got["prim"] != 0
    both = hit_r & hit_g
    same_prim = (ref["prim"] == got["prim"]) & (ref["face"] == got["face"])
    t_close = np.abs(ref["t"] - got["t"]) <= t_rtol * np.abs(ref["t"])
    # A mismatch is excused only as a tie: both hit, at the same distance.  Exact ties (bit-equal
    # t: coincident faces of adjacent boxes, a rect on a box face) are decided by visiting order,
    # and the GPU visits the nearer child first instead of left-then-right; near ties (t within
    # 1e-5 but not equal) must stay rare.
    tie_exact = both & ~same_prim & (ref["t"] == got["t"])
    tie_near = both & ~same_prim & t_close & ~tie_exact
    unexcused = (hit_r != hit_g) | (both & ~same_prim & ~t_close)
    ok = both & same_prim
    stats = {"rays": len(ref), "hits": int(hit_r.sum()), "exact_ties": int(tie_exact.sum()), "near_ties": int(tie_near.sum()),
             "unexcused": int(unexcused.sum())}
    print(label, stats)
    assert stats["unexcused"] == 0, (label, stats, np.flatnonzero(unexcused)[:5])
    assert stats["near_ties"] <= 1e-4 * len(ref) + 2, (label, stats)
    assert np.all(t_close[ok]), (label, "t", np.abs(ref["t"] - got["t"])[ok].max())
    assert np.abs(ref["normal"][ok] - got["normal"][ok]).max() <= n_tol, (label, "normal")
    assert np.array_equal(ref["front"][ok], got["front"][ok]), (label, "front")
    assert np.array_equal(ref["mat"][ok], got["mat"][ok]), (label, "mat")
    p_scale = np.maximum(1.0, np.abs(ref["p"][ok]).max(axis=1))
    assert (np.abs(ref["p"][ok] - got["p"][ok]).max(axis=1) / p_scale).max() <= 10 * t_rtol, (label, "p")
    # (u, v): skip the atan2 seam of Sphere::spherical (u jumps 0 <-> 1)
    # asin(y) is NaN when |y| > 1 by rounding: NaN must then appear on both sides
    assert np.array_equal(np.isnan(ref["u"][ok]), np.isnan(got["u"][ok])) and np.array_equal(np.isnan(ref["v"][ok]), np.isnan(got["v"][ok]))
    du = np.nan_to_num(np.abs(ref["u"][ok] - got["u"][ok]))
    du = np.minimum(du, 1.0 - du)
    dv = np.nan_to_num(np.abs(ref["v"][ok] - got["v"][ok]))
    assert du.max() <= uv_tol and dv.max() <= uv_tol, (label, "uv", du.max(), dv.max())
    return stats


def test_philox_known_answers(vb, ctx):
    """Philox4x32-10 against the Random123 known-answer vectors."""
    L = vb.gpu_lib()
    L.vk_selftest_philox.argtypes = [C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    kat = [([0, 0, 0, 0, 0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
           ([0xffffffff] * 6, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
           ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0], [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1])]
    for inp, want in kat:
        out = (C.c_uint32 * 4)()
        assert L.vk_selftest_philox(ctx._h, (C.c_uint32 * 6)(*inp), out) == 0
        assert list(out) == want


@pytest.mark.parametrize("name,param,width,n", HIT_SCENES, ids=[s[0] for s in HIT_SCENES])
def test_hit_parity_strict(vb, po, ctx, name, param, width, n):
    """Strict math (the reference's op sequence): ids exact, t within 1e-5 (in practice bit-equal)."""
    scene, cam = get_scene(vb, name, param=param)
    o = po.OracleScene(scene)
    ctx.upload(scene)
    rays = ray_batch(vb, o, scene, cam, width, n, seed=11)
    xi = np.random.default_rng(5).random((len(rays), vb.VK_MEDIUM_XI_SLOTS), dtype=np.float32)
    ref = o.intersect(rays, xi)
    got = ctx.intersect(rays, xi, flags=vb.VK_FLAG_STRICT_MATH)
    st = compare_hits(vb, ref, got, 1e-5, 1e-5, 1e-5, name)
    assert st["hits"] > 0.3 * st["rays"]
    ok = (ref["prim"] == got["prim"]) & (ref["prim"] != 0)
    not_medium = ok & ((ref["prim"] >> 28) != vb.VK_T_MEDIUM)
    # distances of deterministic primitives are bit-identical: same IEEE ops in the same order
    assert np.array_equal(ref["t"][not_medium], got["t"][not_medium]), name


@pytest.mark.parametrize("name,param,width,n", HIT_SCENES, ids=[s[0] for s in HIT_SCENES])
def test_hit_parity_fast_math(vb, po, ctx, name, param, width, n):
    """The render build (FMA contraction, reciprocal slab test): same ids except ties, t within 1e-5
    away from grazing hits; reports how often the two builds disagree."""
    scene, cam = get_scene(vb, name, param=param)
    o = po.OracleScene(scene)
    ctx.upload(scene)
    rays = ray_batch(vb, o, scene, cam, width, n, seed=12)
    xi = np.random.default_rng(6).random((len(rays), vb.VK_MEDIUM_XI_SLOTS), dtype=np.float32)
    ref = o.intersect(rays, xi)
    got = ctx.intersect(rays, xi, flags=0)
    hit_r, hit_g = ref["prim"] != 0, got["prim"] != 0
    same = (ref["prim"] == got["prim"]) & (ref["face"] == got["face"])
    with np.errstate(invalid="ignore", divide="ignore"):
        rel = np.abs(ref["t"] - got["t"]) / np.abs(ref["t"])
    # Disagreements between the two builds must be ties (two surfaces at the same distance: touching
    # spheres, box edges) or events at the tmin = 0.001 self-intersection guard -- never a different
    # surface at a different distance.
    tie = hit_r & hit_g & ~same & (rel <= 1e-4)
    # the same surface at a distance that differs by more than 1e-3 relative AND 2e-3 scene units
    # (short hops along the r = 1000 ground sphere lose |oc|^2 - r^2 to cancellation: the relative
    # error of t is large there while the hit point moves by < 1e-4)
    dlen = np.linalg.norm(rays["direction"].astype(np.float64), axis=1)
    with np.errstate(invalid="ignore"):
        far_off = same & hit_r & (rel > 1e-3) & (np.abs(ref["t"].astype(np.float64) - got["t"]) * dlen > 2e-3)
    differ = (hit_r != hit_g) | (hit_r & hit_g & ~same & ~tie) | far_off
    tmin_side = np.minimum(np.where(hit_r, ref["t"], np.inf), np.where(hit_g, got["t"], np.inf)) <= 4e-3
    guard = differ & tmin_side  # one build accepts a self-intersection root at t ~ tmin = 0.001, the other does not
    # Silhouette grazing: half_b^2 - a*c cancels catastrophically in fp32 when a ray grazes a sphere,
    # so the sign of the discriminant -- hit or miss -- depends on FMA contraction.  Verified in
    # fp64: the ray passes within 0.5 % of the radius of one of the two disputed spheres.
    d = scene.desc
    sph = (np.ctypeslib.as_array(C.cast(d.spheres, C.POINTER(C.c_float)), shape=(d.n_spheres, 4)).astype(np.float64)
           if d.n_spheres else np.zeros((1, 4)))
    graze = np.zeros(len(rays), dtype=bool)
    for i in np.flatnonzero(differ & ~tmin_side):
        o64, d64 = rays["origin"][i].astype(np.float64), rays["direction"][i].astype(np.float64)
        for prim in (int(ref["prim"][i]), int(got["prim"][i])):
            if vb.ref_type(prim) == vb.VK_T_SPHERE:
                c, r = sph[vb.ref_index(prim), :3], abs(sph[vb.ref_index(prim), 3])
                oc = c - o64
                rho = np.linalg.norm(oc - d64 * (oc @ d64) / (d64 @ d64))
                graze[i] |= abs(rho - r) <= 5e-3 * r
    other = differ & ~tmin_side & ~graze
    print(f"{name}: fast vs oracle on {len(rays)} rays: {int(tie.sum())} ties, {int(guard.sum())} tmin-guard events, "
          f"{int(graze.sum())} silhouette-grazing events, {int(other.sum())} other")
    for i in np.flatnonzero(other)[:8]:
        print("   other:", hex(ref["prim"][i]), ref["t"][i], hex(got["prim"][i]), got["t"][i], rays["origin"][i], rays["direction"][i])
    # `other` are grazing events the world-space check cannot classify (spheres below an instance)
    # (ties are excused by definition; bowser_demo is built of boxes that touch -- rim on face, arm on body -- and has 0.7 % of them)
    assert tie.mean() <= 1e-2 and guard.mean() <= 5e-4 and graze.mean() <= 1e-3 and other.mean() <= 1e-4, (name, tie.sum(), guard.sum(), graze.sum(), other.sum())
    ok = same & hit_r & ~guard & (rel <= 1e-3)
    # hit POINTS agree to 1e-5 of the scene's coordinate magnitude (t itself loses relative accuracy
    # on short hops along the r = 1000 ground sphere: |oc|^2 - r^2 cancels)
    nodes0 = np.ctypeslib.as_array(C.cast(d.nodes, C.POINTER(C.c_float)), shape=(d.n_nodes, 8))[0]
    scale = float(min(np.abs(np.concatenate([nodes0[0:3], nodes0[4:7]])).max(), 1e4))
    pos_err = np.abs(ref["t"].astype(np.float64) - got["t"])[ok] * dlen[ok]
    print(f"{name}: hit-point error quantiles (99.9 %, max) = {np.quantile(pos_err, 0.999):.2e}, {pos_err.max():.2e} at scene scale {scale:g}")
    assert np.quantile(pos_err, 0.999) <= 1e-5 * scale, (name, np.quantile(pos_err, 0.999), scale)
    assert np.median(rel[ok]) <= 1e-6 and np.quantile(rel[ok], 0.9) <= 1e-5, (name, np.median(rel[ok]), np.quantile(rel[ok], 0.9))
    # Normals.  Flat primitives (rect, box side, medium): the normal is +-e_axis (rotated by an
    # instance), no dependence on t: 1e-5.  Spheres: normal = (p - c) / r (src/hittable.rs:79), so a
    # hit-point difference dp moves it by exactly |dp| / |r|; the bound is that propagation of the
    # (already bounded) hit-point error plus fp32 rounding, not a free tolerance.
    dn = np.linalg.norm(ref["normal"][ok].astype(np.float64) - got["normal"][ok], axis=1)
    dp = np.linalg.norm(ref["p"][ok].astype(np.float64) - got["p"][ok], axis=1)
    ptype = ref["prim"][ok] >> 28
    radius = np.full(int(ok.sum()), np.inf)
    is_s = ptype == vb.VK_T_SPHERE
    radius[is_s] = np.abs(sph[ref["prim"][ok][is_s] & 0x0FFFFFFF, 3])
    if d.n_mspheres:
        msr = np.ctypeslib.as_array(C.cast(d.mspheres, C.POINTER(C.c_float)), shape=(d.n_mspheres, 12))[:, 3].astype(np.float64)
        is_m = ptype == vb.VK_T_MSPHERE
        radius[is_m] = np.abs(msr[ref["prim"][ok][is_m] & 0x0FFFFFFF])
    p_mag = np.maximum(1.0, np.abs(ref["p"][ok]).max(axis=1))
    bound = 1e-5 + (dp + 4e-7 * p_mag) / radius * 2.0
    print(f"{name}: normal error max {dn.max():.2e} (flat primitives: {dn[~np.isfinite(radius)].max() if (~np.isfinite(radius)).any() else 0:.2e})")
    assert np.all(dn <= bound), (name, "normal", float((dn / bound).max()))
    # `front` is the sign of a dot product (src/hittable.rs:23-30): it may flip only where that dot
    # product is within rounding of zero.  For a plain primitive that is a ray tangent to the surface;
    # under a Rotate wrapper the reference dots the OBJECT-space ray with the WORLD-space normal
    # (:618), which vanishes at unrelated angles, so there only the count is bounded.
    flip = np.flatnonzero(ref["front"][ok] != got["front"][ok])
    assert len(flip) <= 1e-4 * ok.sum(), (name, "front", len(flip))
    if d.n_xforms == 0 and len(flip):
        dn64 = rays["direction"][ok][flip].astype(np.float64)
        cosang = np.abs((dn64 * ref["normal"][ok][flip]).sum(axis=1)) / np.linalg.norm(dn64, axis=1)
        assert np.all(cosang <= 1e-2), (name, "front", cosang)


@pytest.mark.parametrize("name,param,width,n", HIT_SCENES, ids=[s[0] for s in HIT_SCENES])
def test_render_build_hits_equal_strict_build_on_a_million_rays(vb, po, ctx, name, param, width, n):
    """The build that renders (FMA contraction, reciprocal slab tests, Boxy by slabs, composed instance transforms) against
    the build whose distances are bit-identical to the oracle's, on 10^6 rays per scene (camera + harvested secondary
    rays), all nine scenes: the same primitive and face except for a small share of ties / grazes / tmin-guard events,
    and the same distance to fp32 rounding where they agree."""
    scene, cam = get_scene(vb, name, param=param)
    o = po.OracleScene(scene)
    ctx.upload(scene)
    rng = np.random.default_rng(77)
    rays = np.concatenate([camera_rays(cam, 600_000, rng), o.harvest_rays(cam, width, scene.height_for(width), 50, 23, 400_000).astype(vb.RAY_DTYPE)])
    xi = rng.random((len(rays), vb.VK_MEDIUM_XI_SLOTS), dtype=np.float32) if scene.desc.n_media else None
    a = ctx.intersect(rays, xi, flags=vb.VK_FLAG_STRICT_MATH)
    b = ctx.intersect(rays, xi, flags=0)
    same = (a["prim"] == b["prim"]) & (a["face"] == b["face"])
    hit = same & (a["prim"] != 0)
    with np.errstate(invalid="ignore", divide="ignore"):
        rel = np.abs(a["t"] - b["t"]) / np.abs(a["t"])
    print(f"{name}: {len(rays)} rays, {int((~same).sum())} differ ({(~same).mean():.2e}), median / 99.9 % relative distance error "
          f"{np.median(rel[hit]):.1e} / {np.quantile(rel[hit], 0.999):.1e}")
    # two surfaces at the same distance (touching boxes: bowser_demo has 0.5 % such rays) may resolve either way
    tie = ~same & (a["prim"] != 0) & (b["prim"] != 0) & (rel <= 1e-4)
    assert (~same & ~tie).mean() <= 3e-3 and tie.mean() <= 1e-2, (name, (~same & ~tie).mean(), tie.mean())
    assert np.median(rel[hit]) <= 1e-6 and np.quantile(rel[hit], 0.99) <= 1e-4
    assert np.array_equal(a["mat"][hit], b["mat"][hit])


def test_intersect_edge_cases(vb, ctx):
    scene, cam = get_scene(vb, "cornell_box")
    ctx.upload(scene)
    assert len(ctx.intersect(np.zeros(0, dtype=vb.RAY_DTYPE))) == 0  # empty batch
    rays = np.zeros(4, dtype=vb.RAY_DTYPE)
    rays["origin"] = [278, 278, -800]
    rays["direction"] = [[0, 0.2, 1], [0, 0, -1], [0, 0, 0], [0, 1e-30, 1e30]]  # hit, miss, zero dir, huge dir
    rays["tmin"], rays["tmax"] = 0.001, INF
    h = ctx.intersect(rays, flags=vb.VK_FLAG_STRICT_MATH)
    assert vb.ref_type(h["prim"][0]) == vb.VK_T_RECT and np.isclose(h["t"][0], 1355.0) and h["prim"][1] == 0
    rays["tmax"] = 100.0  # tmax shorter than the first surface: miss
    assert (ctx.intersect(rays[:1])["prim"] == 0).all()


def zscores(rgb_a, sq_a, n_a, rgb_b, sq_b, n_b):
    """z of the difference of two per-pixel means given their sum-of-squares buffers."""
    var_a = np.maximum(sq_a / n_a - rgb_a.astype(np.float64) ** 2, 0.0) * n_a / (n_a - 1)
    var_b = np.maximum(sq_b / n_b - rgb_b.astype(np.float64) ** 2, 0.0) * n_b / (n_b - 1)
    se = np.sqrt(var_a / n_a + var_b / n_b)
    diff = rgb_a.astype(np.float64) - rgb_b.astype(np.float64)
    z = np.zeros_like(diff)
    # pixels that are constant in both renders (light seen directly, background) have se ~ fp32
    # rounding of sum-of-squares: compare those by value instead
    nz = se > 1e-4 * np.abs(rgb_b) + 1e-12
    z[nz] = diff[nz] / se[nz]
    z[~nz & (np.abs(diff) > 1e-3 * np.abs(rgb_b) + 1e-6)] = np.inf
    return z, nz, diff, se


# Equal spp on both sides: radiance is heavy-tailed, so a low-spp mean is skewed low and its
# variance estimate is correlated with it; with equal spp that skew cancels in the difference.
RENDER_SCENES = [("cornell_box", 0, 80, 512, 512, 100), ("cornell_smoke", 0, 80, 512, 512, 100),
                 ("random_spheres_demo", 0, 128, 256, 256, 50), ("final_scene", 0, 64, 512, 512, 100),
                 ("bowser_demo", 0, 96, 256, 256, 50), ("perlin_demo", 0, 96, 256, 256, 50),
                 ("balls_demo", 0, 96, 256, 256, 50), ("stress_spheres", 64, 96, 128, 128, 50),
                 ("api_surface_demo", 0, 96, 512, 512, 50)]


@pytest.mark.parametrize("name,param,W,spp_o,spp_g,depth", RENDER_SCENES, ids=[s[0] for s in RENDER_SCENES])
def test_image_parity_3_sigma(vb, po, ctx, name, param, W, spp_o, spp_g, depth):
    scene, cam = get_scene(vb, name, param=param)
    H = scene.height_for(W)
    o = po.OracleScene(scene)
    ctx.upload(scene)
    ro, qo, so = o.render(cam, vb.render_params(W, H, spp_o, depth, seed=21), want_sumsq=True)
    rg, qg, sg = ctx.render(cam, vb.render_params(W, H, spp_g, depth, seed=22), want_sumsq=True)
    assert np.isfinite(rg).all()
    assert sg.paths == W * H * spp_g
    # Segments per path must agree (a traversal or termination bug shows up here first).  The GPU
    # stops a path once its weight is exactly zero; the reference keeps tracing such
    # dead segments, so the comparison is against the oracle's count of LIVE rays.
    assert abs(sg.rays / sg.paths - so.rays_live / so.paths) <= 0.02 * so.rays_live / so.paths, (sg.rays / sg.paths, so.rays_live / so.paths)
    z, nz, diff, se = zscores(rg, qg, spp_g, ro, qo, spp_o)
    frac = (np.abs(z[nz]) <= 3.0).mean()
    print(f"{name}: {frac:.4f} of {int(nz.sum())} pixel-channels within 3 sigma, mean z {z[nz].mean():+.3f}, std z {z[nz].std():.3f}, "
          f"image mean ratio {rg.mean() / ro.mean():.4f}")
    assert frac >= 0.985, (name, frac)  # Gaussian expectation 0.9973; heavy-tailed pixels cost a little at these spp
    assert abs(z[nz].mean()) <= 0.1, (name, z[nz].mean())
    assert np.isfinite(z[~nz]).all(), "a pixel is constant in both renders but differs"
    # no spatially coherent bias: z-test of 8x8 tile sums (64x the samples: close to Gaussian)
    th, tw = H // 8, W // 8
    dt = diff[: th * 8, : tw * 8].reshape(th, 8, tw, 8, 3).sum(axis=(1, 3))
    st = np.sqrt((se[: th * 8, : tw * 8] ** 2).reshape(th, 8, tw, 8, 3).sum(axis=(1, 3)))
    zt = dt[st > 0] / st[st > 0]
    ztile = np.where(st > 0, np.abs(dt) / np.where(st > 0, st, 1.0), 0.0).max(axis=2)  # worst channel per tile
    # one firefly (a single sample orders of magnitude above its pixel's mean) can own a tile: at most one such tile
    assert (np.abs(zt) <= 3).mean() >= 0.98 and (ztile > 6.0).sum() <= 1, (name, (np.abs(zt) <= 3).mean(), np.abs(zt).max(), np.argwhere(ztile > 6.0).tolist())
    assert abs(rg.mean() - ro.mean()) <= 0.01 * ro.mean(), (name, rg.mean(), ro.mean())


# SURVEY App. G item 2 at the BASELINE.json configurations' own image sizes: oracle and GPU at equal spp, per pixel and
# channel z = difference of means / combined standard error.  Bars as App. G states them: >= 99 % of the pixel-channels
# within 3 sigma (Gaussian expectation 99.73 %; two oracle renders with different seeds give 99.8-99.9 % at these spp),
# mean z over every 16 x 16 tile within +-0.5 (no spatially coherent bias), image mean within 0.5 %.
CONFIG_SIZED = [("cornell_box", 600, 600, 256, 100), ("cornell_smoke", 600, 600, 128, 100), ("final_scene", 800, 800, 32, 100)]


@pytest.mark.parametrize("name,W,H,spp,depth", CONFIG_SIZED, ids=["config2_cornell_600", "config3_smoke_600", "config4_final_800"])
def test_config_sized_image_parity(vb, po, ctx, name, W, H, spp, depth):
    scene, cam = get_scene(vb, name)
    o = po.OracleScene(scene)
    ctx.upload(scene)
    ro, qo, so = o.render(cam, vb.render_params(W, H, spp, depth, seed=101), want_sumsq=True)
    rg, qg, sg = ctx.render(cam, vb.render_params(W, H, spp, depth, seed=102), want_sumsq=True)
    assert sg.paths == W * H * spp and np.isfinite(rg).all()
    assert abs(sg.rays / sg.paths - so.rays_live / so.paths) <= 0.01 * so.rays_live / so.paths
    z, nz, diff, se = zscores(rg, qg, spp, ro, qo, spp)
    frac = (np.abs(z[nz]) <= 3.0).mean()
    T = 16
    th, tw = H // T, W // T
    zz = np.where(nz, z, 0.0)[: th * T, : tw * T].reshape(th, T, tw, T, 3)
    cnt = nz[: th * T, : tw * T].reshape(th, T, tw, T, 3).sum(axis=(1, 3))
    tile_mean_z = zz.sum(axis=(1, 3))[cnt >= 128] / cnt[cnt >= 128]
    ratio = rg.mean() / ro.mean()
    print(f"{name} {W}x{H}x{spp}: {frac:.4f} of {int(nz.sum())} pixel-channels within 3 sigma, mean z {z[nz].mean():+.4f}, std z {z[nz].std():.3f}, "
          f"max |tile mean z| {np.abs(tile_mean_z).max():.3f} over {tile_mean_z.size} tiles, image mean ratio {ratio:.5f}, "
          f"segments per path {sg.rays / sg.paths:.4f} (oracle, live {so.rays_live / so.paths:.4f})")
    assert frac >= 0.99, (name, frac)
    assert np.abs(tile_mean_z).max() <= 0.5, (name, np.abs(tile_mean_z).max())
    assert abs(z[nz].mean()) <= 0.02, (name, z[nz].mean())
    assert abs(ratio - 1.0) <= 0.005, (name, ratio)
    assert np.isfinite(z[~nz]).all(), "a pixel is constant in both renders but differs"


def test_render_is_deterministic_per_seed(vb, ctx):
    scene, cam = get_scene(vb, "cornell_box")
    ctx.upload(scene)
    p = vb.render_params(160, 160, 64, 100, seed=5)
    a, _, sa = ctx.render(cam, p)
    b, _, sb = ctx.render(cam, p)
    assert np.array_equal(a, b) and sa.rays == sb.rays
    c, _, _ = ctx.render(cam, vb.render_params(160, 160, 64, 100, seed=6))
    assert not np.array_equal(a, c)


def test_spp_slices_combine_like_one_render(vb, ctx):
    """The multi-GPU decomposition on one GPU: per-slice SUM buffers added == the whole render."""
    torch = pytest.importorskip("torch")
    scene, cam = get_scene(vb, "cornell_box")
    ctx.upload(scene)
    W, spp = 128, 96
    n = W * W * 3
    whole, _, _ = ctx.render(cam, vb.render_params(W, W, spp, 100, seed=8))
    acc = torch.zeros(n, dtype=torch.float32, device="cuda:0")
    part = torch.empty_like(acc)
    rays = 0
    for k in range(3):
        st = ctx.render_device(cam, vb.render_params(W, W, spp, 100, seed=8, spp_begin=32 * k, spp_count=32), part.data_ptr())
        acc += part
        rays += st.rays
    out = torch.empty_like(acc)
    torch.cuda.synchronize()  # torch's stream (acc += part) and the context's stream are different streams
    ctx.finalize_device(acc.data_ptr(), out.data_ptr(), n, spp)
    torch.cuda.synchronize()
    assert np.allclose(out.cpu().numpy().reshape(W, W, 3), whole, rtol=1e-5, atol=1e-7)


def test_render_edge_cases(vb, ctx):
    scene, cam = get_scene(vb, "cornell_box")
    ctx.upload(scene)
    # depth 1: only emitters seen directly contribute (one segment, src/main.rs:126)
    rgb, _, st = ctx.render(cam, vb.render_params(64, 64, 16, 1, seed=1))
    assert st.rays == st.paths and rgb.max() == 15.0 and rgb.min() == 0.0
    # ragged image (not a multiple of the 8x4 warp tile) and spp not a multiple of the chunk
    rgb, _, st = ctx.render(cam, vb.render_params(61, 37, 13, 100, seed=1))
    assert rgb.shape == (37, 61, 3) and np.isfinite(rgb).all() and st.paths == 61 * 37 * 13
    for bad in (vb.render_params(1, 64, 4), vb.render_params(64, 64, 0), vb.render_params(64, 64, 4, spp_begin=4),
                vb.render_params(64, 64, 4, spp_begin=2, spp_count=3)):
        with pytest.raises(vb.VecchioError):
            ctx.render(cam, bad)
    fresh = vb.Context(0)
    with pytest.raises(vb.VecchioError) as e:
        fresh.render(cam, vb.render_params(8, 8, 1))
    assert e.value.code == vb.VK_ERR_NO_SCENE
    fresh.close()


def test_golden_cornell_render_matches_published_sample(vb, ctx):
    """GPU Cornell box at the reference's own settings (900^2, 1000 spp, depth 100) against the
    region means of sample/therestofyourlife.png, within +-3 (8-bit) after to_color."""
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "cornell_sample_regions.json")))
    scene, cam = get_scene(vb, "cornell_box")
    ctx.upload(scene)
    rgb, _, st = ctx.render(cam, vb.render_params(900, 900, 1000, 100, seed=1))
    assert st.dropped_samples <= 1e-5 * st.paths
    img = vb.to_color(rgb)[::-1].astype(np.float64)
    for name, r in g["regions"].items():
        x0, x1, y0, y1 = r["box_xyxy"]
        diff = img[y0:y1, x0:x1].mean(axis=(0, 1)) - np.array(r["mean_rgb8"])
        assert np.all(np.abs(diff) <= 3.0), (name, diff)
    nz = np.argwhere(img.sum(axis=2) > 0)
    assert (nz[:, 0].min(), nz[:, 0].max(), nz[:, 1].min(), nz[:, 1].max()) == (22, 879, 21, 878)


def test_golden_final_scene_render_matches_published_sample(vb, ctx):
    """GPU final scene at the reference's own settings (900^2, 1000 spp, depth 100) against the linear region
    means of sample/thenextweek.png (seed-independent objects only: the reference's scene is unseeded).  HEAD's
    integrator must agree within 15 % (20 % on the dark ocean and on the cluster of random spheres); the legacy
    integrator must NOT agree in the fog, i.e. the image does pin HEAD's Isotropic-through-CosinePDF quirk."""
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "thenextweek_sample_regions.json")))
    scene, cam = get_scene(vb, "final_scene")
    ctx.upload(scene)

    def ratios(flags, spp):
        rgb, _, st = ctx.render(cam, vb.render_params(900, 900, spp, 100, seed=1, flags=flags))
        assert st.dropped_samples <= 1e-5 * st.paths
        img = rgb[::-1].astype(np.float64)
        return {name: img[r["box_xyxy"][2]:r["box_xyxy"][3], r["box_xyxy"][0]:r["box_xyxy"][1]].mean(axis=(0, 1)) / np.array(r["mean_linear"])
                for name, r in g["regions"].items()}

    head = ratios(0, 1000)
    for name, ratio in head.items():
        tol = 0.20 if name in ("earth_ocean", "white_cluster") else 0.15
        assert np.all(np.abs(ratio - 1.0) <= tol), (name, ratio)
    legacy = ratios(vb.VK_FLAG_LEGACY_SCATTER, 250)
    assert np.all(legacy["fog_upper_right"] < 0.75) and np.all(legacy["fog_mid_left"] > 1.4), legacy


def test_golden_book1_render_matches_published_sample(vb, ctx):
    """GPU legacy integrator + sky on the book-1 final scene at the published size (1024x576) against the linear
    region means of sample/inoneweekend.png; same tolerances as the oracle's test (they are set by the unseeded
    small spheres, not by noise)."""
    from test_oracle import BOOK1_TOL, book1_region_ratios
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "inoneweekend_sample_regions.json")))
    scene, cam = get_scene(vb, "book1_cover")
    ctx.upload(scene)
    flags = vb.VK_FLAG_LEGACY_SCATTER | vb.VK_FLAG_SKY_BACKGROUND
    rgb, _, st = ctx.render(cam, vb.render_params(1024, 576, 256, 50, seed=1, flags=flags))
    # (refract()'s sqrt of a rounding-negative number makes a NaN ray, the reference's own Q7: such a sample is dropped
    # exactly as main.rs:192 drops it; which of 1.5e8 samples hit that case depends on FMA contraction)
    assert st.dropped_samples <= 1e-7 * st.paths
    for name, ratio in book1_region_ratios(g, rgb[::-1].astype(np.float64)).items():
        assert np.all(np.abs(ratio - 1.0) <= BOOK1_TOL[name]), (name, ratio)


def test_full_size_cornell_properties(vb, ctx):
    """BASELINE.json config 2 at full size (600x600, 1000 spp): size-independent properties."""
    scene, cam = get_scene(vb, "cornell_box")
    ctx.upload(scene)
    rgb, _, st = ctx.render(cam, vb.render_params(600, 600, 1000, 100, seed=1))
    assert st.paths == 360_000_000 and np.isfinite(rgb).all() and rgb.min() >= 0.0
    assert 2.65 <= st.rays / st.paths <= 2.9  # oracle: 3.07 segments per path, 2.78 of them with non-zero weight
    assert st.dropped_samples <= 1e-5 * st.paths
    # mirror symmetry of radiance is broken only by the wall colours: red wall redder, green greener
    left, right = rgb[250:350, 20:60], rgb[250:350, 540:580]  # image-left is +x (green wall)
    assert left[..., 1].mean() > 2 * left[..., 0].mean() and right[..., 0].mean() > 4 * right[..., 1].mean()
    # linearity in spp slices: the first half's mean equals the whole within noise
    half, _, _ = ctx.render(cam, vb.render_params(600, 600, 500, 100, seed=1))
    assert abs(half.mean() - rgb.mean()) <= 0.005 * rgb.mean()


# ---------------------------------------------------------------------------------------------------
# Wavefront variant (generate -> extend -> material-sorted shade -> accumulate as separate kernels)
# ---------------------------------------------------------------------------------------------------
WF_SCENES = [("cornell_box", 0, 96, 64, 100), ("cornell_smoke", 0, 64, 64, 100), ("final_scene", 0, 64, 32, 100),
             ("random_spheres_demo", 0, 96, 32, 50), ("bowser_demo", 0, 64, 32, 50), ("stress_spheres", 40, 96, 8, 50)]


@pytest.mark.parametrize("variant", [2, 3, 4, 5], ids=["wavefront", "staged", "warpq", "stepq"])
@pytest.mark.parametrize("name,param,W,spp,depth", WF_SCENES, ids=[s[0] for s in WF_SCENES])
def test_variants_equal_megakernel(vb, ctx, name, param, W, spp, depth, variant):
    """Every variant keys Philox by (pixel, global sample, bounce) and adds a finished sample to its pixel's
    integer (fixed-point) accumulators, so the order in which paths run does not matter: with the strict
    build (no FMA contraction, so the shared device functions round identically in all kernels) the images
    are bit-identical, and so are the segment and dropped-sample counts.  The fast build agrees to fp32
    rounding.  (The sums of squares are double-precision atomics: equal to the last bits only.)"""
    scene, cam = get_scene(vb, name, param=param)
    H = scene.height_for(W)
    ctx.upload(scene)
    for flags, exact in ((vb.VK_FLAG_STRICT_MATH, True), (0, False)):
        pm = vb.render_params(W, H, spp, depth, seed=31, variant=vb.VK_VARIANT_MEGAKERNEL, flags=flags)
        pw = vb.render_params(W, H, spp, depth, seed=31, variant=variant, flags=flags)
        a, qa, sa = ctx.render(cam, pm, want_sumsq=True)
        b, qb, sb = ctx.render(cam, pw, want_sumsq=True)
        # (step queues are the BVH traversal: on a flat-program scene the request runs the flat warp-queue kernel)
        ran = vb.VK_VARIANT_WARPQ if (variant == vb.VK_VARIANT_STEPQ and sb.node_visits == 0) else variant
        assert sb.variant == ran and sa.variant == vb.VK_VARIANT_MEGAKERNEL
        assert (sb.launches > 3 or variant != vb.VK_VARIANT_WAVEFRONT) and sa.paths == sb.paths
        if exact:
            assert np.array_equal(a, b) and np.allclose(qa, qb, rtol=1e-6, atol=0), (name, np.abs(a - b).max())
            assert (sa.rays, sa.dropped_samples) == (sb.rays, sb.dropped_samples)
        else:
            # contraction differs between the two kernels: a path can take another branch at a rounding
            # boundary, so compare statistically tight instead of bitwise
            assert abs(sa.rays - sb.rays) <= 2e-3 * sa.rays
            close = np.isclose(a, b, rtol=1e-3, atol=1e-5)
            assert close.mean() >= 0.97, (name, close.mean())
            assert abs(a.mean() - b.mean()) <= 2e-3 * a.mean()


def test_wavefront_image_parity_with_oracle(vb, po, ctx):
    scene, cam = get_scene(vb, "cornell_box")
    o = po.OracleScene(scene)
    ctx.upload(scene)
    W, spp = 80, 512
    ro, qo, so = o.render(cam, vb.render_params(W, W, spp, 100, seed=41), want_sumsq=True)
    rg, qg, sg = ctx.render(cam, vb.render_params(W, W, spp, 100, seed=42, variant=vb.VK_VARIANT_WAVEFRONT), want_sumsq=True)
    assert sg.variant == vb.VK_VARIANT_WAVEFRONT
    z, nz, diff, se = zscores(rg, qg, spp, ro, qo, spp)
    frac = (np.abs(z[nz]) <= 3.0).mean()
    print(f"wavefront cornell: {frac:.4f} within 3 sigma, mean z {z[nz].mean():+.3f}")
    assert frac >= 0.985 and abs(z[nz].mean()) <= 0.1
    assert abs(sg.rays / sg.paths - so.rays_live / so.paths) <= 0.02 * so.rays_live / so.paths


@pytest.mark.parametrize("wf", [2, 3, 4], ids=["wavefront", "staged", "warpq"])
def test_wavefront_edge_cases(vb, ctx, wf):
    scene, cam = get_scene(vb, "cornell_box")
    ctx.upload(scene)
    # fewer units than pool slots, ragged image, spp not a multiple of the sample block, depth 1
    rgb, _, st = ctx.render(cam, vb.render_params(61, 37, 13, 100, seed=1, variant=wf))
    ref, _, sr = ctx.render(cam, vb.render_params(61, 37, 13, 100, seed=1, variant=vb.VK_VARIANT_MEGAKERNEL))
    assert rgb.shape == (37, 61, 3) and st.paths == 61 * 37 * 13 and abs(rgb.mean() - ref.mean()) <= 5e-3 * ref.mean()
    rgb, _, st = ctx.render(cam, vb.render_params(64, 64, 16, 1, seed=1, variant=wf))
    assert st.rays == st.paths and rgb.max() == 15.0 and rgb.min() == 0.0
    # an spp slice (the multi-GPU unit of work) through the wavefront
    a, _, _ = ctx.render(cam, vb.render_params(64, 64, 32, 100, seed=3, variant=wf, flags=vb.VK_FLAG_STRICT_MATH))
    b, _, _ = ctx.render(cam, vb.render_params(64, 64, 32, 100, seed=3, flags=vb.VK_FLAG_STRICT_MATH))
    assert np.array_equal(a, b)


def test_rgb8_frame_is_to_color_of_the_float_frame(vb, ctx):
    """vk_render_rgb8 = the frame loop's output stage on the device (src/main.rs:201-214): to_color per
    channel, rows top-down.  Same seed -> same samples, so it must equal the host-side to_color of
    vk_render's floats bit for bit; and a turntable keeps the scene resident between cameras."""
    scene = vb.Scene("random_spheres_demo")  # RotatingCamera: 671 frames (src/scene.rs:254-281)
    ctx.upload(scene)
    W, H = 96, scene.height_for(96)
    frames = []
    for _ in range(3):
        cam = scene.next_camera()
        p = vb.render_params(W, H, 8, 50, seed=4)
        rgb, _, _ = ctx.render(cam, p)
        rgb8, st = ctx.render_rgb8(cam, p)
        assert rgb8.shape == (H, W, 3) and st.paths == W * H * 8
        assert np.array_equal(rgb8, vb.to_color(rgb)[::-1])
        frames.append(rgb8)
    assert not np.array_equal(frames[0], frames[1])  # the camera moved
    # known answers of to_color: 0.25 -> 128, 1.0 -> 255, 0 -> 0
    assert list(vb.to_color(np.array([0.25, 1.0, 0.0, -1.0, np.nan], dtype=np.float32))) == [128, 255, 0, 0, 0]


# ---------------------------------------------------------------------------------------------------
# Legacy book-1/2 integrator (VK_FLAG_LEGACY_SCATTER) and sky background
# ---------------------------------------------------------------------------------------------------
LEGACY_CASES = [("random_spheres_cover", 96, 128, 50, True), ("cornell_box", 64, 512, 50, False), ("final_scene", 48, 128, 50, False)]


@pytest.mark.parametrize("variant", [1, 4, 5], ids=["megakernel", "warpq", "stepq"])
@pytest.mark.parametrize("name,W,spp,depth,sky", LEGACY_CASES, ids=[c[0] for c in LEGACY_CASES])
def test_legacy_integrator_image_parity(vb, po, ctx, name, W, spp, depth, sky, variant):
    """The legacy integrator runs in the lane megakernel and in the warp-queue kernel (flat program and BVH)."""
    scene, cam = get_scene(vb, name)
    H = scene.height_for(W)
    o = po.OracleScene(scene)
    ctx.upload(scene)
    flags = vb.VK_FLAG_LEGACY_SCATTER | (vb.VK_FLAG_SKY_BACKGROUND if sky else 0)
    ro, qo, so = o.render(cam, vb.render_params(W, H, spp, depth, seed=51, flags=flags), want_sumsq=True)
    rg, qg, sg = ctx.render(cam, vb.render_params(W, H, spp, depth, seed=52, flags=flags, variant=variant), want_sumsq=True)
    assert sg.variant == (vb.VK_VARIANT_WARPQ if (variant == vb.VK_VARIANT_STEPQ and sg.node_visits == 0) else variant) and np.isfinite(rg).all()
    assert abs(sg.rays / sg.paths - so.rays / so.paths) <= 0.02 * so.rays / so.paths, (sg.rays / sg.paths, so.rays / so.paths)
    z, nz, diff, se = zscores(rg, qg, spp, ro, qo, spp)
    frac = (np.abs(z[nz]) <= 3.0).mean()
    print(f"legacy {name}: {frac:.4f} within 3 sigma, mean z {z[nz].mean():+.3f}, image mean ratio {rg.mean() / ro.mean():.4f}")
    assert frac >= 0.98 and abs(z[nz].mean()) <= 0.1, (name, frac, z[nz].mean())
    assert abs(rg.mean() - ro.mean()) <= 0.02 * ro.mean()


def test_legacy_and_light_list_rules(vb, ctx):
    cover, cam = get_scene(vb, "random_spheres_cover")
    ctx.upload(cover)  # no lights: accepted at upload
    with pytest.raises(vb.VecchioError) as e:  # HEAD's integrator panics on an empty light list (src/hittable.rs:431)
        ctx.render(cam, vb.render_params(32, 18, 4, 10))
    assert e.value.code == vb.VK_ERR_INVALID
    rgb, _, st = ctx.render(cam, vb.render_params(32, 18, 4, 10, flags=vb.VK_FLAG_LEGACY_SCATTER | vb.VK_FLAG_SKY_BACKGROUND))
    assert rgb.mean() > 0.3  # sky lit
    # the sky alone: depth 1 from above the scene straight up is the top of the gradient
    balls, cam2 = get_scene(vb, "api_surface_demo")  # holds a SpecDiffuse
    ctx.upload(balls)
    with pytest.raises(vb.VecchioError) as e:
        ctx.render(cam2, vb.render_params(32, 18, 4, 10, flags=vb.VK_FLAG_LEGACY_SCATTER))
    assert e.value.code == vb.VK_ERR_UNSUPPORTED
    # sky background also works under HEAD's integrator (any variant): brighter than the black default
    a, _, _ = ctx.render(cam2, vb.render_params(64, 36, 16, 20, seed=1, flags=vb.VK_FLAG_SKY_BACKGROUND, variant=vb.VK_VARIANT_WAVEFRONT))
    b, _, _ = ctx.render(cam2, vb.render_params(64, 36, 16, 20, seed=1, variant=vb.VK_VARIANT_WAVEFRONT))
    assert a.mean() > 1.5 * b.mean()


# ---------------------------------------------------------------------------------------------------
# BASELINE.json configs 1, 3, 4, 5 at their full image sizes: size-independent properties
# (config 2 is test_full_size_cornell_properties).  Samples per pixel are reduced for 4 and 5, and said so.
# ---------------------------------------------------------------------------------------------------
def test_full_size_config1_random_spheres(vb, po, ctx):
    """Config 1 is small enough for the oracle at full size (400x225, 16 spp, depth 50): image statistics
    of both renders must agree, not just a crop."""
    scene, cam = get_scene(vb, "random_spheres_demo")
    ctx.upload(scene)
    o = po.OracleScene(scene)
    rg, qg, sg = ctx.render(cam, vb.render_params(400, 225, 16, 50, seed=61), want_sumsq=True)
    ro, qo, so = o.render(cam, vb.render_params(400, 225, 16, 50, seed=62), want_sumsq=True)
    assert sg.paths == 400 * 225 * 16 == so.paths and np.isfinite(rg).all()
    assert abs(sg.rays / sg.paths - so.rays_live / so.paths) <= 0.02 * so.rays_live / so.paths
    z, nz, _, _ = zscores(rg, qg, 16, ro, qo, 16)
    assert (np.abs(z[nz]) <= 3).mean() >= 0.97  # 16 spp: heavy-tailed pixels cost a little more than at 256+
    assert abs(rg.mean() - ro.mean()) <= 0.03 * ro.mean()


def test_full_size_config3_cornell_smoke(vb, ctx):
    scene, cam = get_scene(vb, "cornell_smoke")
    ctx.upload(scene)
    rgb, _, st = ctx.render(cam, vb.render_params(600, 600, 2000, 100, seed=1))
    assert st.paths == 720_000_000 and np.isfinite(rgb).all() and rgb.min() >= 0.0
    assert 2.1 <= st.rays / st.paths <= 2.4 and st.dropped_samples <= 1e-5 * st.paths
    # the black smoke box (image right: x ~ 265..430 world = image-left mirrored) is darker than the white one
    img = rgb[::-1]  # top-down
    lum = img.mean(axis=2)
    # the two media sit where the Cornell boxes are: tall box region (dark smoke) vs short box region (white smoke)
    tall, short = lum[250:400, 170:270].mean(), lum[400:520, 330:440].mean()
    assert short > 1.3 * tall, (tall, short)
    half, _, _ = ctx.render(cam, vb.render_params(600, 600, 1000, 100, seed=1))
    assert abs(half.mean() - rgb.mean()) <= 0.005 * rgb.mean()  # linearity in spp slices


def test_full_size_config4_final_scene(vb, po, ctx):
    """800x800 at 256 spp instead of the config's 10 000 (the properties do not depend on spp)."""
    scene, cam = get_scene(vb, "final_scene")
    ctx.upload(scene)
    rgb, _, st = ctx.render(cam, vb.render_params(800, 800, 256, 100, seed=1))
    assert st.paths == 800 * 800 * 256 and np.isfinite(rgb).all() and rgb.min() >= 0.0
    _, _, so = po.OracleScene(scene).render(cam, vb.render_params(100, 100, 16, 100, seed=2))
    assert abs(st.rays / st.paths - so.rays_live / so.paths) <= 0.03 * so.rays_live / so.paths  # segments per path: size independent
    assert st.dropped_samples <= 1e-4 * st.paths
    again, _, st2 = ctx.render(cam, vb.render_params(800, 800, 256, 100, seed=1))
    assert np.array_equal(rgb, again) and st2.rays == st.rays  # deterministic per seed


def test_full_size_config5_million_spheres(vb, po, ctx):
    """10^6 spheres in one reference-built BVH at 3840x2160, 4 spp instead of 256."""
    scene = vb.Scene("stress_spheres", seed=1, param=1000)
    cam = scene.next_camera()
    assert scene.census()["spheres"] == 1_000_000
    ctx.upload(scene)
    rgb, _, st = ctx.render(cam, vb.render_params(3840, 2160, 4, 50, seed=1))
    assert st.paths == 3840 * 2160 * 4 and np.isfinite(rgb).all() and rgb.min() >= 0.0
    assert 2.7 <= st.rays / st.paths <= 2.9 and st.dropped_samples <= 1e-3 * st.paths
    # Segments per path and image mean against the oracle at the oracle's size, with the RENDER (fast) build.
    # This scene is where FMA contraction in Sphere::hit / Ray::at once changed the statistics (|oc|^2 - r^2
    # cancels at |oc| ~ 300, r = 0.2: 4.4 % more segments, 1.5 % darker); both are now never contracted.
    small, _, ss = ctx.render(cam, vb.render_params(192, 108, 16, 50, seed=3))
    ro, _, so = po.OracleScene(scene).render(cam, vb.render_params(192, 108, 16, 50, seed=2))
    assert abs(ss.rays / ss.paths - so.rays_live / so.paths) <= 0.01 * so.rays_live / so.paths, (ss.rays / ss.paths, so.rays_live / so.paths)
    assert abs(small.mean() - ro.mean()) <= 0.01 * ro.mean(), (small.mean(), ro.mean())
    # spp slices of the same seed add up to the whole (the multi-GPU decomposition) at this size too
    a, _, _ = ctx.render(cam, vb.render_params(3840, 2160, 4, 50, seed=1, spp_begin=0, spp_count=2))
    b, _, _ = ctx.render(cam, vb.render_params(3840, 2160, 4, 50, seed=1, spp_begin=2, spp_count=2))
    assert np.allclose(a + b, rgb, rtol=1e-5, atol=1e-7)
    scene.close()


STAT_SCENES = [("cornell_box", 0), ("cornell_smoke", 0), ("random_spheres_demo", 0), ("final_scene", 0), ("bowser_demo", 0),
               ("perlin_demo", 0), ("balls_demo", 0), ("api_surface_demo", 0), ("stress_spheres", 300)]


@pytest.mark.parametrize("name,param", STAT_SCENES, ids=[s[0] for s in STAT_SCENES])
def test_render_build_has_the_statistics_of_the_strict_build(vb, ctx, name, param):
    """The render build (FMA contraction, approximate reciprocal / rsqrt) against the strict build (the
    reference's operation sequence) with the same seed: segments per path within 0.5 %, image mean
    within 1 %, dropped samples of the same order.  Guards against contraction changing which
    near-degenerate intersections are found (it did, for spheres far from the origin)."""
    scene, cam = get_scene(vb, name, param=param)
    ctx.upload(scene)
    W = 256
    H = scene.height_for(W)
    a, _, sa = ctx.render(cam, vb.render_params(W, H, 16, 50, seed=71))
    b, _, sb = ctx.render(cam, vb.render_params(W, H, 16, 50, seed=71, flags=vb.VK_FLAG_STRICT_MATH))
    ra, rb = sa.rays / sa.paths, sb.rays / sb.paths
    print(f"{name}: segments per path fast {ra:.4f} strict {rb:.4f}; mean fast {a.mean():.5f} strict {b.mean():.5f}; dropped {sa.dropped_samples} / {sb.dropped_samples}")
    assert abs(ra - rb) <= 5e-3 * rb, (name, ra, rb)
    assert abs(a.mean() - b.mean()) <= 1e-2 * b.mean(), (name, a.mean(), b.mean())
    assert abs(int(sa.dropped_samples) - int(sb.dropped_samples)) <= 0.5 * sb.dropped_samples + 50


def test_accumulation_does_not_depend_on_who_renders_what(vb, ctx):
    """Finished samples go into 64-bit fixed-point accumulators (integer atomics), so a frame is a function of
    (seed, sample range, size) alone: the four variants agree bit for bit (strict build), a re-run agrees bit for
    bit, and the per-pixel SUMS of two disjoint sample ranges add up to the sums of the whole range exactly as
    integers -- checked here through the means, whose only rounding is the final fp32 conversion."""
    scene, cam = get_scene(vb, "cornell_box")
    ctx.upload(scene)
    p = lambda v, **kw: vb.render_params(96, 96, 48, 100, seed=9, variant=v, flags=vb.VK_FLAG_STRICT_MATH, **kw)  # noqa: E731
    full, _, s0 = ctx.render(cam, p(vb.VK_VARIANT_WARPQ))
    again, _, s1 = ctx.render(cam, p(vb.VK_VARIANT_WARPQ))
    assert s0.rays == s1.rays and np.array_equal(full, again)
    for v in (vb.VK_VARIANT_MEGAKERNEL, vb.VK_VARIANT_WAVEFRONT, vb.VK_VARIANT_STAGED):
        other, _, _ = ctx.render(cam, p(v))
        assert np.array_equal(full, other), v
    lo, _, _ = ctx.render(cam, p(vb.VK_VARIANT_WARPQ, spp_begin=0, spp_count=20))
    hi, _, _ = ctx.render(cam, p(vb.VK_VARIANT_MEGAKERNEL, spp_begin=20, spp_count=28))
    assert np.allclose(lo.astype(np.float64) + hi, full, rtol=3e-7, atol=1e-9)


def test_hybrid_program_renders_the_same_image(vb, ctx, monkeypatch):
    """The hybrid flat program (typed batches for the mixed top of the scene, its homogeneous subtrees walked as BVHs;
    what VK_VARIANT_WARPQ runs on such a scene, and with VECCHIO_HYBRID_ALL=1 every other kernel too) is the same
    function as the BVH: identical hits on a ray batch and a bit-identical strict-build image."""
    scene, cam = get_scene(vb, "final_scene")
    ctx.upload(scene)
    p = vb.render_params(64, 64, 16, 50, seed=5, flags=vb.VK_FLAG_STRICT_MATH, variant=vb.VK_VARIANT_MEGAKERNEL)
    a, _, sa = ctx.render(cam, p)
    rays = camera_rays(cam, 20000, np.random.default_rng(3))
    ha = ctx.intersect(rays, flags=vb.VK_FLAG_STRICT_MATH)
    pw = vb.render_params(64, 64, 16, 50, seed=5, flags=vb.VK_FLAG_STRICT_MATH, variant=vb.VK_VARIANT_WARPQ)
    w, _, sw = ctx.render(cam, pw)  # k_warpq_hybrid
    monkeypatch.setenv("VECCHIO_HYBRID_ALL", "1")
    b, _, sb = ctx.render(cam, p)  # the lane megakernel over the hybrid program
    hb = ctx.intersect(rays, flags=vb.VK_FLAG_STRICT_MATH)
    monkeypatch.delenv("VECCHIO_HYBRID_ALL")
    assert sb.node_visits < sa.node_visits and sw.node_visits < sa.node_visits  # the flat top replaces the upper nodes
    same = ha["prim"] == hb["prim"]
    assert same.mean() >= 0.999 and np.array_equal(ha["t"][same], hb["t"][same])  # exact ties may resolve differently
    assert np.array_equal(a, b) or np.isclose(a, b, rtol=1e-5, atol=1e-6).mean() > 0.999
    assert sw.variant == vb.VK_VARIANT_WARPQ and (np.array_equal(a, w) or np.isclose(a, w, rtol=1e-5, atol=1e-6).mean() > 0.999)


@pytest.mark.parametrize("name", ["cornell_box", "final_scene", "random_spheres_demo"])
def test_hit_parity_strict_ten_million_rays(vb, po, ctx, name):
    """SURVEY App. G / build-plan step 3: >= 10^7 rays per scene through the oracle's `world.hit` and
    `vk_intersect` (strict build): ids exact except exact-t ties, t bit-identical, normals / p / uv / front /
    material within 1e-5.  8M camera rays + 2M secondary rays harvested from oracle paths (unnormalised
    cosine, light and specular directions)."""
    scene, cam = get_scene(vb, name)
    o = po.OracleScene(scene)
    ctx.upload(scene)
    rng = np.random.default_rng(101)
    cam_rays = camera_rays(cam, 8_000_000, rng)
    sec = o.harvest_rays(cam, 512, scene.height_for(512), 50, 17, 2_000_000)
    rays = np.concatenate([cam_rays, sec.astype(cam_rays.dtype)])
    assert len(rays) >= 9_500_000  # the harvest may stop a little short on scenes with few bounces
    xi = None
    if scene.desc.n_media:
        xi = rng.random((len(rays), vb.VK_MEDIUM_XI_SLOTS), dtype=np.float32)
    ref = o.intersect(rays, xi)
    got = ctx.intersect(rays, xi, flags=vb.VK_FLAG_STRICT_MATH)
    st = compare_hits(vb, ref, got, 1e-5, 1e-5, 1e-5, f"{name} x {len(rays)}")
    ok = (ref["prim"] == got["prim"]) & (ref["prim"] != 0) & ((ref["prim"] >> 28) != vb.VK_T_MEDIUM)
    assert np.array_equal(ref["t"][ok], got["t"][ok])
    assert st["hits"] > 0.3 * st["rays"]
