#!/bin/bash
# usage: gpu_scale.sh N...   -- bench.py at each N on this box (torchrun for N > 1)
mkdir -p gpurun_out
for n in "$@"; do
  if [ "$n" = 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  fi
  echo "N=$n rc=$?"; tail -c 2500 gpurun_out/scale_n$n.json | python -c "
import sys, json
for line in sys.stdin:
    try: d = json.loads(line)
    except Exception: continue
    print({k: d[k] for k in ('value', 'n_gpus', 'ms_per_step', 'mrays_per_s', 'gpu_launches')}, 'e2e', d['e2e']['value'], 'roofline', d['roofline'] and d['roofline']['frac'])
"
  tail -3 gpurun_out/scale_n$n.err
done
