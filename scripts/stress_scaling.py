import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vecchio_b200 as vb
ctx = vb.Context(0)
for param in (32, 100, 316, 1000):
    t = time.time(); s = vb.Scene("stress_spheres", seed=1, param=param); tb = time.time() - t
    cam = s.next_camera(); t = time.time(); ctx.upload(s); tu = time.time() - t
    for rep in range(2):
        rgb, _, st = ctx.render(cam, vb.render_params(960, 540, 4, 50, seed=1 + rep))
    print(f"param {param}: build {tb:.2f}s upload {tu:.3f}s kernels {st.ms_kernels:.1f} ms, {st.rays/st.ms_kernels/1e3:.1f} Mrays/s, nodes/ray {st.node_visits/st.rays:.1f} prims/ray {st.prim_tests/st.rays:.1f} rays/path {st.rays/st.paths:.2f}", flush=True)
