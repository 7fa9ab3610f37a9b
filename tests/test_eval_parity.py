"""Per-function parity of the SHADING half of the path: `vk_eval_batch` (device functions behind
Material::scatter_with_pdf / scattering_pdf / emitted, Texture::value, the light list's pdf_value / random and the
mixture estimator of ray_color, src/main.rs:131-149) against the oracle's `orc_eval_batch`, record by record, with
the SAME supplied variates.  vk_intersect covers the geometry half the same way (test_gpu_parity.py).

Strict build: the direction formulas are the reference's operation for operation -> 1e-6; weights divide in a
different association (iterative prefix product) -> 1e-5.  Fast build (FMA contraction, rsqrt): 2e-5 / 1e-4.
Rejection-sampled pieces (random_in_unit_sphere: Metal with fuzz, legacy Isotropic) are the same law but not the
same points: their direction is checked against the law's support, everything else exactly."""
import numpy as np
import pytest

from conftest import get_scene

EVAL_SCENES = ["cornell_box", "cornell_smoke", "final_scene", "random_spheres_demo", "api_surface_demo", "bowser_demo", "perlin_demo"]


def unit(v):
    return v / np.linalg.norm(v, axis=-1, keepdims=True)


def scene_extent(desc):
    lo, hi = np.full(3, np.inf), np.full(3, -np.inf)
    for i in range(desc.n_nodes):
        lo = np.minimum(lo, np.array(desc.nodes[i].bb_min[:]))
        hi = np.maximum(hi, np.array(desc.nodes[i].bb_max[:]))
    big = 2000.0  # the ground sphere (r = 1000) and the r = 5000 fog boundary would dominate: sample the part a camera sees
    return np.maximum(lo, -big), np.minimum(hi, big)


def bounce_records(vb, desc, rng, per_material, op):
    lo, hi = scene_extent(desc)
    n = desc.n_materials * per_material
    r = np.zeros(n, dtype=vb.EVAL_DTYPE)
    r["op"] = op
    r["index"] = np.repeat(np.arange(desc.n_materials, dtype=np.uint32), per_material)
    r["p"] = rng.uniform(lo, hi, (n, 3)).astype(np.float32)
    nrm = unit(rng.normal(size=(n, 3)))
    nrm[::4] = np.eye(3)[rng.integers(0, 3, len(nrm[::4]))] * rng.choice([-1.0, 1.0], (len(nrm[::4]), 1))  # axis-aligned ones (rects)
    d = rng.normal(size=(n, 3)) * rng.uniform(0.2, 30.0, (n, 1))  # unnormalised directions (Q5)
    flip = np.sum(d * nrm, axis=1) > 0  # a HitRec's normal faces the ray (set_face_normal) ...
    nrm[flip] *= -1.0
    r["normal"] = nrm.astype(np.float32)
    r["ray_d"] = d.astype(np.float32)
    r["ray_o"] = (r["p"] - 0.5 * r["ray_d"]).astype(np.float32)
    r["ray_time"] = rng.uniform(0, 1, n).astype(np.float32)
    r["t"] = 0.5
    r["u"] = rng.uniform(0, 1, n).astype(np.float32)
    r["v"] = rng.uniform(0, 1, n).astype(np.float32)
    r["front"] = rng.integers(0, 2, n)  # ... while `front` is whatever the wrappers left (FlipFace, Q9)
    r["xi"] = rng.integers(0, 2 ** 32, (n, 5), dtype=np.uint64).astype(np.uint32)
    return r


def material_types(desc):
    """per material: (type, resolved type behind SpecDiffuse is not needed), fuzz"""
    t = np.array([desc.materials[i].type for i in range(desc.n_materials)])
    prm = np.array([desc.materials[i].param for i in range(desc.n_materials)], dtype=np.float32)
    return t, prm


def compare_bounces(vb, desc, ref, got, tol_dir, tol_w, legacy=False):
    """Returns counts for the caller's coverage asserts."""
    mt, prm = material_types(desc)
    typ = mt[ref["index"]]
    # SpecDiffuse picks specular / diffuse with xi[4]; its children decide what is comparable
    spec_child = np.array([desc.materials[desc.materials[i].tex].type if mt[i] == vb.VK_M_SPECDIFFUSE else -1 for i in range(desc.n_materials)])
    diff_child = np.array([desc.materials[desc.materials[i].aux].type if mt[i] == vb.VK_M_SPECDIFFUSE else -1 for i in range(desc.n_materials)])
    pct = prm[ref["index"]]
    u4 = (ref["xi"][:, 4] >> 8).astype(np.float32) * np.float32(2.0 ** -24)
    eff = np.where(typ == vb.VK_M_SPECDIFFUSE, np.where(u4 < pct, spec_child[ref["index"]], diff_child[ref["index"]]), typ)
    eff_fuzz = np.where(typ == vb.VK_M_SPECDIFFUSE,
                        np.array([desc.materials[desc.materials[i].tex].param if mt[i] == vb.VK_M_SPECDIFFUSE else 0.0 for i in range(desc.n_materials)], dtype=np.float32)[ref["index"]],
                        prm[ref["index"]])
    rejection = ((eff == vb.VK_M_METAL) & (eff_fuzz != 0.0)) | (legacy & (eff == vb.VK_M_ISOTROPIC))
    # emission and validity
    assert np.array_equal(ref["L"], got["L"]) or np.allclose(ref["L"], got["L"], rtol=1e-6, atol=0)
    ref_w0 = np.all(ref["beta"] == 0.0, axis=1)
    ref_bad = ~np.all(np.isfinite(ref["beta"]), axis=1)
    assert np.array_equal(got["valid"] == 0, ref_bad | (ref["valid"] == 0)), "validity (non-finite weight) differs"
    ok = ~ref_bad & (ref["valid"] != 0)
    # the GPU ends a path whose weight became 0 (nothing further can contribute); the reference recurses with weight 0.
    # A cosine that is zero to rounding (|cos| < 1e-5) may be positive on one side and negative on the other.
    alive_ref = (ref["alive"] != 0) & ~ref_w0
    alive_got = (got["alive"] != 0) & ~np.all(got["beta"] == 0.0, axis=1)
    und = unit(ref["out_d"].astype(np.float64) + 1e-300)
    grazing = (np.abs(np.sum(und * ref["normal"], axis=1)) < 1e-4) | (np.abs(ref["beta"]).max(axis=1) < 1e-4) | (np.abs(got["beta"]).max(axis=1) < 1e-4)
    differ = ok & ~rejection & (alive_got != alive_ref)  # (a rejection-sampled direction may be absorbed on one side only: legacy Metal)
    assert not (differ & ~grazing).any(), ("continue / end differs", int((differ & ~grazing).sum()), eff[differ & ~grazing][:8].tolist())
    assert differ.sum() <= 1e-3 * len(ref), ("too many grazing disagreements", int(differ.sum()))
    cont = ok & alive_ref & alive_got
    det = cont & ~rejection
    # A weight is albedo * cos / pdf and the direction itself carries ~1e-6 of rounding (sincos, the association of
    # 2 pi r1): the weight's error is that over the cosine, so it is bounded relative to the weight only away from grazing
    cosn = np.abs(np.sum(und * ref["normal"], axis=1))
    tol_i = (tol_w + 3.0 * tol_dir / np.maximum(cosn, 1e-3))[:, None]
    # a NoiseTexture albedo is sin(scale z + 10 turb) with seven octaves of fp32 sums behind an argument of order 10^2: its
    # absolute error (see the texture records below: 2e-5 strict, 2e-3 render build) times spdf / pdf <= 2 enters the weight
    def noisy(ti, depth=0):
        t = desc.textures[ti]
        return t.type == vb.VK_TEX_NOISE or (t.type == vb.VK_TEX_CHECKER and depth < 8 and (noisy(t.w[0], depth + 1) or noisy(t.w[1], depth + 1)))
    mat_noisy = np.array([desc.materials[i].type not in (vb.VK_M_DIELECTRIC, vb.VK_M_SPECDIFFUSE) and noisy(desc.materials[i].tex) for i in range(desc.n_materials)])
    tex_atol = np.where(mat_noisy[ref["index"]], 4.0 * (2e-5 if tol_dir <= 1e-6 else 2e-3), 0.0)[:, None]
    werr = np.abs(got["beta"].astype(np.float64) - ref["beta"]) - tol_i * np.abs(ref["beta"]) - tex_atol
    wbad = det[:, None] & (werr > 1e-7)
    assert not wbad.any(), ("weights differ", int(wbad.sum()), float(werr[wbad].max()), eff[wbad.any(axis=1)][:8].tolist(),
                            got["beta"][wbad.any(axis=1)][:4].tolist(), ref["beta"][wbad.any(axis=1)][:4].tolist(), cosn[wbad.any(axis=1)][:4].tolist())
    # directions: error relative to the magnitude of what was added up (a light-sampled direction is point - origin,
    # legacy Lambertian is normal + unit vector: both can cancel)
    u0 = (ref["xi"][:, 0] >> 8).astype(np.float32) * np.float32(2.0 ** -24)
    light_branch = (not legacy) & (u0 < 0.5) & np.isin(eff, (vb.VK_M_LAMBERTIAN, vb.VK_M_ISOTROPIC))
    scale = np.abs(ref["out_d"]).max(axis=1, keepdims=True) + 1.0 + (np.abs(ref["p"]).max(axis=1) * light_branch)[:, None]
    derr = np.abs(got["out_d"].astype(np.float64) - ref["out_d"])
    dbad = det[:, None] & (derr > tol_dir * scale)
    assert not dbad.any(), ("directions differ", int(dbad.sum()), float(derr[dbad].max()), eff[dbad.any(axis=1)][:8].tolist())
    assert np.array_equal(got["out_o"][cont], ref["out_o"][cont]) and np.array_equal(got["out_time"][cont], ref["out_time"][cont])
    # rejection-sampled directions: inside the law's support, same attenuation
    rj = cont & rejection
    if rj.any():
        assert np.allclose(got["beta"][rj], ref["beta"][rj], rtol=tol_w)
        m = rj & (eff == vb.VK_M_METAL)
        if m.any():  # reflect(unit(d), n) + fuzz * (point in the unit ball)
            ud = unit(ref["ray_d"][m].astype(np.float64))
            nn = ref["normal"][m].astype(np.float64)
            refl = ud - 2.0 * np.sum(ud * nn, axis=1, keepdims=True) * nn
            for side in (got, ref):
                assert np.all(np.linalg.norm(side["out_d"][m] - refl, axis=1) <= eff_fuzz[m] * (1 + 1e-4) + 1e-5)
        i = rj & (eff == vb.VK_M_ISOTROPIC)
        if i.any():
            for side in (got, ref):
                assert np.all(np.linalg.norm(side["out_d"][i], axis=1) <= 1 + 1e-5)
    return {"records": len(ref), "continue": int(cont.sum()), "deterministic": int(det.sum()), "rejection": int(rj.sum()),
            "ended": int((ok & ~alive_ref).sum()), "dropped": int((~ok).sum()), "types": sorted(set(eff.tolist()))}


def texture_records(vb, desc, rng, per_texture):
    lo, hi = scene_extent(desc)
    n = desc.n_textures * per_texture
    r = np.zeros(n, dtype=vb.EVAL_DTYPE)
    r["op"] = vb.VK_EVAL_TEXTURE
    r["index"] = np.repeat(np.arange(desc.n_textures, dtype=np.uint32), per_texture)
    r["p"] = rng.uniform(np.maximum(lo, -50), np.minimum(hi, 50), (n, 3)).astype(np.float32)
    r["u"] = rng.uniform(-0.1, 1.1, n).astype(np.float32)  # ImageTexture clamps (src/material.rs:283-284)
    r["v"] = rng.uniform(-0.1, 1.1, n).astype(np.float32)
    return r


def checker_margin(p):
    """|sin(10x) sin(10y) sin(10z)|: a Checker (and anything under it) may flip across its zero set within rounding"""
    return np.abs(np.sin(10.0 * p[:, 0].astype(np.float64)) * np.sin(10.0 * p[:, 1].astype(np.float64)) * np.sin(10.0 * p[:, 2].astype(np.float64)))


def light_records(vb, desc, rng, n_each, oracle):
    lo, hi = scene_extent(desc)
    nl = desc.n_lights
    r = np.zeros(nl * n_each, dtype=vb.EVAL_DTYPE)
    r["op"] = vb.VK_EVAL_LIGHT_RANDOM
    r["index"] = np.repeat(np.arange(nl, dtype=np.uint32), n_each)
    r["p"] = rng.uniform(lo, hi, (len(r), 3)).astype(np.float32)
    r["xi"] = rng.integers(0, 2 ** 32, (len(r), 5), dtype=np.uint64).astype(np.uint32)
    # pdf records: directions the oracle sampled towards the lights (non-zero pdfs) plus random ones (mostly zero)
    sampled = oracle.eval_batch(r)
    q = np.zeros(2 * len(r), dtype=vb.EVAL_DTYPE)
    q["op"] = vb.VK_EVAL_LIGHTS_PDF
    q["p"] = np.concatenate([r["p"], r["p"]])
    q["dir"] = np.concatenate([sampled["out_d"], rng.normal(size=(len(r), 3)).astype(np.float32)])
    return r, q


def run_eval_parity(vb, scene, oracle, evaluate, strict):
    """`evaluate(recs) -> recs` is the side under test (the GPU hook; the oracle itself in the CPU self-check)."""
    desc = scene.desc
    rng = np.random.default_rng(2024)
    # render build: __sincosf on [0, 2 pi) (absolute error grows past pi), rsqrt, FMA contraction
    tol_dir, tol_w = (1e-6, 1e-5) if strict else (1e-4, 1e-4)
    out = {}
    if desc.n_lights:
        recs = bounce_records(vb, desc, rng, 4000, vb.VK_EVAL_BOUNCE)
        out["head"] = compare_bounces(vb, desc, oracle.eval_batch(recs), evaluate(recs), tol_dir, tol_w)
    has_specdiffuse = any(desc.materials[i].type == vb.VK_M_SPECDIFFUSE for i in range(desc.n_materials))
    if not has_specdiffuse:
        recs = bounce_records(vb, desc, rng, 2000, vb.VK_EVAL_BOUNCE_LEGACY)
        out["legacy"] = compare_bounces(vb, desc, oracle.eval_batch(recs), evaluate(recs), tol_dir, tol_w, legacy=True)
    recs = texture_records(vb, desc, rng, 3000)
    ref, got = oracle.eval_batch(recs), evaluate(recs)
    safe = checker_margin(recs["p"]) > 1e-3
    tt = np.array([desc.textures[i].type for i in range(desc.n_textures)])[recs["index"]]
    noise = tt == vb.VK_TEX_NOISE  # sin(scale z + 10 turb): seven octaves of fp32 sums, argument of order 10^2
    assert np.allclose(got["beta"][safe & ~noise], ref["beta"][safe & ~noise], rtol=1e-6, atol=0)
    assert np.allclose(got["beta"][safe & noise], ref["beta"][safe & noise], rtol=0, atol=2e-5 if strict else 2e-3)
    out["textures"] = {"records": len(recs), "types": sorted(set(tt.tolist()))}
    if desc.n_lights:
        r, q = light_records(vb, desc, rng, 3000, oracle)
        ref, got = oracle.eval_batch(r), evaluate(r)
        scale = np.maximum(np.abs(ref["out_d"]).max(axis=1, keepdims=True), 1.0)
        assert np.all(np.abs(got["out_d"] - ref["out_d"]) <= (1e-6 if strict else 1e-5) * scale)
        ref, got = oracle.eval_batch(q), evaluate(q)
        both = (ref["value"] > 0) & (got["value"] > 0)
        # a direction that grazes a light's edge may be in on one side and out on the other
        assert (both | ((ref["value"] == 0) & (got["value"] == 0))).mean() > 0.999
        # a sphere light's pdf is 1 / (2 pi (1 - cos_max)) with cos_max = sqrt(1 - r^2 / d^2) (src/hittable.rs:104-111): for a
        # small far light 1 - cos_max cancels, an fp32 rounding of cos_max is a relative error of eps * 2 pi * pdf
        rel = np.abs(got["value"][both].astype(np.float64) / ref["value"][both] - 1.0)
        assert np.all(rel <= (1e-5 if strict else 1e-4) + (4e-7 if strict else 2e-6) * ref["value"][both]), float(rel.max())
        out["lights"] = {"random": len(r), "pdf": len(q), "pdf_nonzero": int(both.sum())}
    return out


@pytest.mark.parametrize("name", EVAL_SCENES)
def test_eval_harness_self_check(vb, po, name):
    """CPU: the harness itself (record generation, comparison logic, coverage) with the oracle on both sides --
    and the scripted variates must make the oracle reproducible call to call."""
    scene, _ = get_scene(vb, name)
    o = po.OracleScene(scene)
    out = run_eval_parity(vb, scene, o, o.eval_batch, strict=True)
    if "head" in out:
        assert out["head"]["continue"] > 0.3 * out["head"]["records"]
    assert out["textures"]["records"] > 0


@pytest.mark.gpu
@pytest.mark.parametrize("strict", [True, False], ids=["strict", "fast"])
@pytest.mark.parametrize("name", EVAL_SCENES)
def test_shading_functions_match_the_oracle_record_by_record(vb, po, ctx, name, strict):
    scene, _ = get_scene(vb, name)
    o = po.OracleScene(scene)
    ctx.upload(scene)
    flags = vb.VK_FLAG_STRICT_MATH if strict else 0
    out = run_eval_parity(vb, scene, o, lambda r: ctx.eval_batch(r, flags), strict)
    print(name, "strict" if strict else "fast", out)
    if "head" in out:
        assert out["head"]["continue"] > 0.3 * out["head"]["records"]


@pytest.mark.gpu
def test_eval_covers_every_material_texture_and_light_kind(vb, po, ctx):
    """Across the scenes above: all six materials, all four textures, Rect / Sphere / Boxy lights."""
    mats, texs, lights = set(), set(), set()
    for name in EVAL_SCENES:
        scene, _ = get_scene(vb, name)
        d = scene.desc
        mats |= {d.materials[i].type for i in range(d.n_materials)}
        texs |= {d.textures[i].type for i in range(d.n_textures)}
        lights |= {vb.ref_type(d.lights[i]) for i in range(d.n_lights)}
    assert mats == set(range(6)) and texs == set(range(4)) and {vb.VK_T_RECT, vb.VK_T_SPHERE, vb.VK_T_BOX} <= lights


@pytest.mark.gpu
def test_eval_rejects_bad_records(vb, ctx):
    scene, _ = get_scene(vb, "cornell_box")
    ctx.upload(scene)
    r = np.zeros(1, dtype=vb.EVAL_DTYPE)
    r["op"] = 9
    with pytest.raises(vb.VecchioError):
        ctx.eval_batch(r)
    r["op"], r["index"] = vb.VK_EVAL_BOUNCE, 10 ** 6
    with pytest.raises(vb.VecchioError):
        ctx.eval_batch(r)
    assert len(ctx.eval_batch(np.zeros(0, dtype=vb.EVAL_DTYPE))) == 0
