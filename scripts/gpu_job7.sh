#!/bin/bash
mkdir -p gpurun_out
python scripts/render_once.py cornell 200 3 > gpurun_out/plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_staged -s 1 -c 1 -f -o gpurun_out/prof_r1_staged_v2 \
    python scripts/render_once.py cornell 200 3 > gpurun_out/ncu_full3.log 2>&1; echo "full rc=$?"
cat gpurun_out/plain3.log
