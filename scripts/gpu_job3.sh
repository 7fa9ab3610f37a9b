#!/bin/bash
# wavefront evidence: launch list (time per kernel) and one full capture of extend + shade mid-run
mkdir -p gpurun_out
python scripts/render_once.py cornell 100 2 > gpurun_out/plain_wf.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 200 --csv --log-file gpurun_out/launches_wf.csv \
    python scripts/render_once.py cornell 100 2 > gpurun_out/ncu_wf_launch.log 2>&1; echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_wf_ -s 300 -c 2 -f -o gpurun_out/prof_r1_wf \
    python scripts/render_once.py cornell 100 2 > gpurun_out/ncu_wf_full.log 2>&1; echo "full rc=$?"
cat gpurun_out/plain_wf.log
