"""One render through the C ABI (for ncu): python scripts/render_once.py [config] [spp] [variant] [width]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vecchio_b200 as vb
from bench import CONFIGS

cfg = sys.argv[1] if len(sys.argv) > 1 else "cornell"
scene_name, param, W, H, spp, depth, _ = CONFIGS[cfg]
spp = int(sys.argv[2]) if len(sys.argv) > 2 else spp
variant = int(sys.argv[3]) if len(sys.argv) > 3 else 0
if len(sys.argv) > 4:
    W = int(sys.argv[4])
scene = vb.Scene(scene_name, seed=1, param=param)
if len(sys.argv) > 4:
    H = scene.height_for(W)
cam = scene.next_camera()
ctx = vb.Context(0)
ctx.upload(scene)
for rep in range(2):
    rgb, _, st = ctx.render(cam, vb.render_params(W, H, spp, depth, seed=1 + rep, variant=variant))
print(f"{cfg} {W}x{H}x{spp}: {st.ms_kernels:.2f} ms kernels, {st.paths / st.ms_kernels / 1e3:.1f} Mpaths/s, "
      f"{st.rays / st.ms_kernels / 1e3:.1f} Mrays/s, rays/path {st.rays / st.paths:.3f}, nodes/ray {st.node_visits / st.rays:.1f}, prims/ray {st.prim_tests / st.rays:.1f}, mean {rgb.mean():.5f}")
