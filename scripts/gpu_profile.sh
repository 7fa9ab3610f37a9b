#!/bin/bash
# ncu evidence for the bench step: launch list of bench.py, full capture of the dominant kernel at the bench config
#   gpurun --timeout 900 -- 'bash scripts/gpu_profile.sh'
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/ncu_launch.log 2>&1; echo "launch list rc=$?"
python scripts/render_once.py cornell 1000 0 > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_warpq -s 1 -c 1 -f -o gpurun_out/prof_r2_cornell_warpq \
    python scripts/render_once.py cornell 1000 0 > gpurun_out/ncu_full.log 2>&1; echo "full rc=$?"
cat gpurun_out/plain2.log
