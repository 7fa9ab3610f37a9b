"""Per-CTA end time and ray count of the staged kernel.  usage: cta_spread.py config spp"""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vecchio_b200 as vb
from bench import CONFIGS
cfg, spp = sys.argv[1], int(sys.argv[2])
scene_name, param, W, H, _, depth, _ = CONFIGS[cfg]
scene = vb.Scene(scene_name, seed=1, param=param); cam = scene.next_camera()
ctx = vb.Context(0); ctx.upload(scene)
L = vb.gpu_lib(); L.vk_debug_counters.argtypes = [C.c_void_p, C.POINTER(C.c_ulonglong)]; L.vk_debug_ctas.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
for rep in range(2):
    _, _, st = ctx.render(cam, vb.render_params(W, H, spp, depth, seed=1 + rep, variant=3))
out = (C.c_ulonglong * 8)(); L.vk_debug_counters(ctx._h, out)
n = ctx.device_info()["sm_count"] * 3
d = np.zeros((n, 4), dtype=np.uint64); L.vk_debug_ctas(ctx._h, d.ctypes.data, n)
t = (d[:, 0].astype(np.int64) - int(out[5])) / 1e6; rays = d[:, 1].astype(np.float64); sm = d[:, 2].astype(int)
print(f"{cfg} spp {spp}: kernel {st.ms_kernels:.2f} ms; CTA end min/median/max {t.min():.2f} {np.median(t):.2f} {t.max():.2f} ms; rays per CTA min/mean/max {rays.min():.0f} {rays.mean():.0f} {rays.max():.0f} (std {rays.std()/rays.mean()*100:.2f} %)")
it = (d[:, 3] & np.uint64(0xFFFFFFFF)).astype(int); sp = (d[:, 3] >> np.uint64(32)).astype(int)
print(f"iterations per CTA min/median/max {it.min()} {int(np.median(it))} {it.max()}; sparse (<N/8 live) iterations min/median/max {sp.min()} {int(np.median(sp))} {sp.max()}")
print("corr(end time, rays) =", np.corrcoef(t, rays)[0, 1])
rate = rays / t
print("rays/ms per CTA min/median/max", rate.min(), np.median(rate), rate.max())
bysm = {}
for i in range(n): bysm.setdefault(sm[i], []).append(t[i])
ends = np.array([max(v) for v in bysm.values()]); cnt = np.array([len(v) for v in bysm.values()])
print("SMs used", len(bysm), "CTAs per SM min/max", cnt.min(), cnt.max(), "; per-SM end min/median/max", ends.min(), np.median(ends), ends.max())
order = np.argsort(t)
print("slowest 8 CTAs (cta, sm, end, rays):", [(int(i), int(sm[i]), round(float(t[i]), 2), int(rays[i])) for i in order[-8:]])
print("fastest 8 CTAs:", [(int(i), int(sm[i]), round(float(t[i]), 2), int(rays[i])) for i in order[:8]])
