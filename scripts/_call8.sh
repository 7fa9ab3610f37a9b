set -x
nvidia-smi --query-gpu=index,name --format=csv | head -3
python -m pytest tests/test_zz_frames_gpu.py -m gpu -q -k "two_gpus or multi_context" > gpurun_out/r2_pytest_2gpu.log 2>&1; tail -2 gpurun_out/r2_pytest_2gpu.log
python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/r2_scale_n1.json 2> gpurun_out/r2_scale_n1.err; cut -c1-200 gpurun_out/r2_scale_n1.json
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/r2_scale_n$n.json 2> gpurun_out/r2_scale_n$n.err; cut -c1-200 gpurun_out/r2_scale_n$n.json
done
python - > gpurun_out/r2_multi_8gpu.log 2>&1 <<'PY'
import time, numpy as np, vecchio_b200 as vb
s = vb.Scene("cornell_box"); cam = s.next_camera()
one = vb.Context(0); one.upload(s)
p = vb.render_params(600, 600, 1000, 100, seed=1)
a, _, sa = one.render(cam, p)
for n in (2, 4, 8):
    m = vb.MultiContext(list(range(n))); m.upload(s)
    for rep in range(3):
        t1 = time.perf_counter(); b, _, sb = m.render(cam, p); t2 = time.perf_counter()
    print(f"vk_multi_render on {n} GPUs: {1e3*(t2-t1):.2f} ms wall (slowest device's kernels {sb.ms_kernels:.2f}, total {sb.ms_total:.2f}); one GPU kernels {sa.ms_kernels:.2f}; identical frames: {np.array_equal(a, b)}", flush=True)
    del m
PY
cat gpurun_out/r2_multi_8gpu.log
