set -x
python scripts/render_once.py cornell_smoke 64 0 > gpurun_out/plain_s.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_warpq -s 1 -c 1 -f -o gpurun_out/prof_r2_smoke_warpq \
    python scripts/render_once.py cornell_smoke 64 0 > gpurun_out/ncu_full_s.log 2>&1; echo "full rc=$?"
python scripts/render_once.py final_scene 16 0 > gpurun_out/plain_f.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_megakernel -s 1 -c 1 -f -o gpurun_out/prof_r2_final_scene_megakernel \
    python scripts/render_once.py final_scene 16 0 > gpurun_out/ncu_full_f.log 2>&1; echo "full rc=$?"
cat gpurun_out/plain_s.log gpurun_out/plain_f.log
