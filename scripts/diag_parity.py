"""Diagnostic (GPU box): image-parity statistics per scene at a higher oracle spp, and a per-step
timing probe of the device loop."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vecchio_b200 as vb
from oracle import pyoracle as po

def z(rg, qg, ng, ro, qo, no):
    va = np.maximum(qg / ng - rg.astype(np.float64) ** 2, 0) * ng / (ng - 1)
    vo = np.maximum(qo / no - ro.astype(np.float64) ** 2, 0) * no / (no - 1)
    se = np.sqrt(va / ng + vo / no)
    d = rg.astype(np.float64) - ro
    ok = se > 1e-4 * np.abs(ro) + 1e-12
    return d[ok] / se[ok], ok

ctx = vb.Context(0)
for name, W, so_, sg_, depth in [("final_scene", 64, 512, 4096, 100), ("bowser_demo", 64, 512, 4096, 50), ("perlin_demo", 64, 512, 4096, 50),
                                  ("random_spheres_demo", 96, 256, 2048, 50), ("cornell_smoke", 64, 512, 4096, 100)]:
    s = vb.Scene(name); cam = s.next_camera(); H = s.height_for(W)
    o = po.OracleScene(s); ctx.upload(s)
    t = time.time(); ro, qo, st = o.render(cam, vb.render_params(W, H, so_, depth, seed=21), True); to = time.time() - t
    rg, qg, sg = ctx.render(cam, vb.render_params(W, H, sg_, depth, seed=22), True)
    zz, ok = z(rg, qg, sg_, ro, qo, so_)
    print(f"{name:22s} oracle {to:5.1f}s frac3 {np.mean(np.abs(zz)<=3):.4f} frac2 {np.mean(np.abs(zz)<=2):.4f} meanz {zz.mean():+.3f} stdz {zz.std():.3f} "
          f"mean_ratio {rg.mean()/ro.mean():.4f} rays/path gpu {sg.rays/sg.paths:.3f} live {st.rays_live/st.paths:.3f} dropped {sg.dropped_samples}/{st.dropped_samples}", flush=True)

import torch
s = vb.Scene("cornell_box"); cam = s.next_camera(); ctx.upload(s)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
n = 600 * 600 * 3
d_sum = torch.empty(n, device="cuda"); d_rgb = torch.empty(n, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
evs = [torch.cuda.Event(enable_timing=True) for _ in range(20)]
for mode in ("noflush", "flush"):
    k = 0
    for i in range(6):
        evs[k].record(stream); k += 1
        if mode == "flush": flush.zero_()
        evs[k].record(stream); k += 1
        ctx.render_device(cam, vb.render_params(600, 600, 1000, 100, seed=i + 1), d_sum.data_ptr(), want_stats=False)
        ctx.finalize_device(d_sum.data_ptr(), d_rgb.data_ptr(), n, 1000)
        evs[k].record(stream); k += 1
    torch.cuda.synchronize()
    print(mode, [f"{evs[3*i].elapsed_time(evs[3*i+1]):.2f}+{evs[3*i+1].elapsed_time(evs[3*i+2]):.1f}" for i in range(6)], flush=True)
st = ctx.flush_stats()
print("rays/path", st.rays / st.paths)
