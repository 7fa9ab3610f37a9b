set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
python scripts/diag_furnace.py > gpurun_out/r2_diag_furnace.log 2>&1
python - > gpurun_out/r2_ab.log 2>&1 <<'PY'
import time, numpy as np
import vecchio_b200 as vb
ctx = vb.Context(0)
for name, W, H, spp in (("cornell_box", 600, 600, 1000), ("cornell_smoke", 600, 600, 500)):
    scene = vb.Scene(name); cam = scene.next_camera(); ctx.upload(scene)
    ref = None
    for v in (3, 4, 1):
        for rep in range(3):
            rgb, _, st = ctx.render(cam, vb.render_params(W, H, spp, 100, seed=1, variant=v))
        print(name, "variant", v, "ms_kernels %.3f" % st.ms_kernels, "rays/path %.4f" % (st.rays / st.paths), "dropped", st.dropped_samples,
              "mean %.6f" % rgb.mean(), "Mpaths/s %.1f" % (st.paths / st.ms_kernels / 1e3), flush=True)
        if ref is None: ref = rgb
        else: print("   max rel diff of image mean vs variant 3: %.3e" % abs(rgb.mean() / ref.mean() - 1))
PY
python -m pytest tests -m gpu -x -q -rxX > gpurun_out/r2_pytest_gpu_0.log 2>&1
python bench.py > gpurun_out/r2_bench_0.log 2> gpurun_out/r2_bench_0.err
tail -3 gpurun_out/r2_diag_furnace.log gpurun_out/r2_pytest_gpu_0.log gpurun_out/r2_bench_0.log; cat gpurun_out/r2_ab.log
