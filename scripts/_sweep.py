"""A/B sweep helper (one process per library: VECCHIO_GPU_LIB is read at import)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vecchio_b200 as vb
tag = sys.argv[1]
jobs = [a.split(":") for a in sys.argv[2:]]  # scene:W:H:spp:depth:variant
ctx = vb.Context(0)
scenes = {}
for name, W, H, spp, depth, variant in jobs:
    if name not in scenes:
        nm, _, prm = name.partition("@")
        s = vb.Scene(nm, param=int(prm or 0)); scenes[name] = (s, s.next_camera())
    s, cam = scenes[name]
    ctx.upload(s)
    best = None
    for rep in range(3):
        rgb, _, st = ctx.render(cam, vb.render_params(int(W), int(H), int(spp), int(depth), seed=1, variant=int(variant)))
        best = st.ms_kernels if best is None else min(best, st.ms_kernels)
    print(f"{tag:8s} {name:22s} {W}x{H}x{spp} variant {variant} ran {st.variant}: {best:8.3f} ms  {st.paths / best / 1e3:9.1f} Mpaths/s  rays/path {st.rays / st.paths:.4f} "
          f"nodes/ray {st.node_visits / max(st.rays, 1):.2f} dropped {st.dropped_samples} mean {rgb.mean():.6f}", flush=True)
