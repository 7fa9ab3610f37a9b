mkdir -p gpurun_out
python scripts/render_once.py final_scene 64 0 > gpurun_out/plain_fs.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_megakernel -s 1 -c 1 -f -o gpurun_out/prof_r1_final_scene_v2 \
    python scripts/render_once.py final_scene 64 0 > gpurun_out/ncu_fs.log 2>&1; echo "full rc=$?"
cat gpurun_out/plain_fs.log
