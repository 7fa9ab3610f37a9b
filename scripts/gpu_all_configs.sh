#!/bin/bash
# usage: gpu_all_configs.sh N   -- bench.py for every BASELINE.json config on N GPUs of this box.
# At N = 1 configs 4 and 5 run an eighth of their spp (the per-GPU share of the 8-GPU run) and say so.
N=$1; mkdir -p gpurun_out
run() { # config, extra args
  local c=$1; shift
  if [ "$N" = 1 ]; then timeout 900 python bench.py --gpus 1 --steps 2 --warmup 3 --config $c "$@" > gpurun_out/cfg_${c}_n$N.json 2> gpurun_out/cfg_${c}_n$N.err
  else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 2 --warmup 3 --config $c "$@" > gpurun_out/cfg_${c}_n$N.json 2> gpurun_out/cfg_${c}_n$N.err; fi
  echo "$c N=$N rc=$?"; tail -c 4000 gpurun_out/cfg_${c}_n$N.json | python -c "
import sys, json
for line in sys.stdin:
    try: d = json.loads(line)
    except Exception: continue
    print('  ', d['config']['workload'][:70], '|', round(d['value'],1), 'Mpaths/s', round(d['mrays_per_s'],1), 'Mrays/s', round(d['ms_per_step'],2), 'ms | e2e', round(d['e2e']['value'],1), '| frac', d['roofline'] and round(d['roofline']['frac'],4), '| cpu', d.get('cpu_baseline') and round(d['cpu_baseline']['value'],3))
"; tail -2 gpurun_out/cfg_${c}_n$N.err | cut -c1-200
}
if [ "$N" = 1 ]; then
  run random_spheres; run cornell; run cornell_smoke; run final_scene --spp 1250; run stress_1m --spp 32
else
  run cornell_smoke; run final_scene; run stress_1m
fi
