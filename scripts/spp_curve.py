"""Kernel time against spp for one config and variant (fixed cost vs per-sample cost), with the
CTA end-time spread of the staged kernel.  usage: spp_curve.py config variant"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vecchio_b200 as vb
from bench import CONFIGS
cfg, variant = sys.argv[1], int(sys.argv[2])
scene_name, param, W, H, spp, depth, _ = CONFIGS[cfg]
scene = vb.Scene(scene_name, seed=1, param=param); cam = scene.next_camera()
ctx = vb.Context(0); ctx.upload(scene)
L = vb.gpu_lib(); L.vk_debug_counters.argtypes = [C.c_void_p, C.POINTER(C.c_ulonglong)]
for s in (8, 16, 32, 64, 125, 250, 500, 1000):
    for rep in range(3):
        _, _, st = ctx.render(cam, vb.render_params(W, H, s, depth, seed=1 + rep, variant=variant))
    out = (C.c_ulonglong * 8)(); L.vk_debug_counters(ctx._h, out)
    spread = f"first CTA end {(out[6]-out[5])/1e6:.2f} ms, last {(out[7]-out[5])/1e6:.2f} ms" if variant == 3 else ""
    print(f"{cfg} v{variant} spp {s:5d}: {st.ms_kernels:8.3f} ms kernels, total {st.ms_total:8.3f} ms, {st.paths/st.ms_kernels/1e3:8.1f} Mpaths/s  {spread}", flush=True)
