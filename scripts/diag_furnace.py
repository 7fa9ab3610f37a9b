"""Diagnose the white-medium furnace on the GPU: which assert of tests/test_zz_furnace_gpu.py trips, per variant / math build."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import vecchio_b200 as vb

E = np.array([0.8, 0.6, 0.4])
scene = vb.Scene("furnace_demo", param=3)
cam = scene.next_camera()
ctx = vb.Context(0)
ctx.upload(scene)
W, spp = 128, 2048
yy, xx = np.mgrid[0:W, 0:W]
disc = ((yy - 63.5) ** 2 + (xx - 63.5) ** 2) < 24 ** 2
for variant in (0, 1, 2, 3, 4):
    for flags in (0, vb.VK_FLAG_STRICT_MATH, vb.VK_FLAG_FORCE_BVH, vb.VK_FLAG_STRICT_MATH | vb.VK_FLAG_FORCE_BVH):
        for seed in (1, 2):
            try:
                rgb, sq, st = ctx.render(cam, vb.render_params(W, W, spp, 100, seed=seed, variant=variant, flags=flags), want_sumsq=True)
            except Exception as e:
                print("variant", variant, "flags", flags, "ERR", e); continue
            rgb = rgb.astype(np.float64)
            mean = rgb[disc].mean(axis=0)
            sigma = np.sqrt((sq[disc] / spp - rgb[disc] ** 2).mean(axis=0) / spp / disc.sum())
            print(f"variant {variant} ran {st.variant} flags {flags} seed {seed}: dropped {st.dropped_samples} rays/path {st.rays/st.paths:.4f} "
                  f"mean/E {mean/E} z {(mean-E)/sigma} corner/E {rgb[:8,:8].mean(axis=(0,1))/E}", flush=True)
for variant in (0, 1, 4):
    for flags in (vb.VK_FLAG_LEGACY_SCATTER, vb.VK_FLAG_LEGACY_SCATTER | vb.VK_FLAG_STRICT_MATH):
        legacy, _, st = ctx.render(cam, vb.render_params(W, W, 16, 100, seed=1, variant=variant, flags=flags))
        rel = np.abs(legacy / E - 1)
        bad = np.argwhere(rel > 1e-5)
        print(f"legacy variant {variant} flags {flags}: dropped {st.dropped_samples} max rel err {rel.max():.3e} n bad {len(bad)} first {bad[:5].tolist()} "
              f"vals {[legacy[tuple(b[:2])].tolist() for b in bad[:3]]}", flush=True)
