set -x
nvidia-smi --query-gpu=index,name --format=csv
python -m pytest tests/test_zz_frames_gpu.py -m gpu -q -rs > gpurun_out/r2_pytest_2gpu.log 2>&1; tail -4 gpurun_out/r2_pytest_2gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_scale_n2.json 2> gpurun_out/r2_scale_n2.err; cat gpurun_out/r2_scale_n2.json | cut -c1-400
python - > gpurun_out/r2_multi_2gpu.log 2>&1 <<'PY'
import time, numpy as np, vecchio_b200 as vb
s = vb.Scene("cornell_box"); cam = s.next_camera()
one = vb.Context(0); one.upload(s)
m = vb.MultiContext([0, 1]); m.upload(s)
p = vb.render_params(600, 600, 1000, 100, seed=1)
for rep in range(3):
    t0 = time.perf_counter(); a, _, sa = one.render(cam, p); t1 = time.perf_counter(); b, _, sb = m.render(cam, p); t2 = time.perf_counter()
    print(f"one GPU vk_render {1e3*(t1-t0):.2f} ms (kernels {sa.ms_kernels:.2f}); vk_multi_render on 2 GPUs {1e3*(t2-t1):.2f} ms (slowest device's kernels {sb.ms_kernels:.2f}, total {sb.ms_total:.2f}); identical frames: {np.array_equal(a, b)}; rays {sa.rays} / {sb.rays}", flush=True)
PY
cat gpurun_out/r2_multi_2gpu.log
vecchio_b200/lib/vecchio_gpu_render --scene 1 --width 600 --spp 1000 --gpus 2 --out-dir gpurun_out --frames 1 2>&1 | tail -2
