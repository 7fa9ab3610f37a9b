# ncu evidence of the final build: bench launch list, full captures of the Cornell warp-queue kernel and the final-scene lane megakernel
set -x
bash scripts/gpu_profile.sh
python scripts/render_once.py final_scene 16 0 > gpurun_out/plain_f.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_megakernel -s 1 -c 1 -f -o gpurun_out/prof_r2_final_scene_megakernel \
    python scripts/render_once.py final_scene 16 0 > gpurun_out/ncu_full_f.log 2>&1; echo "full rc=$?"
cat gpurun_out/plain_f.log
