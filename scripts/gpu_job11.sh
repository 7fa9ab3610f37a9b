#!/bin/bash
mkdir -p gpurun_out
python scripts/render_once.py stress_1m 1 1 > gpurun_out/plain_st.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_megakernel -s 1 -c 1 -f -o gpurun_out/prof_r1_stress_dyn \
    python scripts/render_once.py stress_1m 1 1 > gpurun_out/ncu_st.log 2>&1; echo "full rc=$?"
cat gpurun_out/plain_st.log
