set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
python -m pytest tests -m gpu -q -rxXs > gpurun_out/r2_pytest_gpu_3.log 2>&1
tail -5 gpurun_out/r2_pytest_gpu_3.log
FLAT="cornell_box:600:600:1000:100:4 cornell_smoke:600:600:500:100:4"
rm -f gpurun_out/r2_sweep_3.log
for tag in base fr0 k3 n144 rg8 rg24 mk2 mn176; do
  VECCHIO_GPU_LIB=build/libvk_$tag.so python scripts/_sweep.py $tag $FLAT >> gpurun_out/r2_sweep_3.log 2>&1
done
python scripts/_sweep.py default cornell_box:600:600:1000:100:3 cornell_smoke:600:600:500:100:3 cornell_box:600:600:1000:100:4 cornell_smoke:600:600:500:100:4 \
   cornell_box:600:600:125:100:4 cornell_box:600:600:125:100:3 \
   final_scene:800:800:64:100:1 final_scene:800:800:64:100:4 random_spheres_demo:400:225:256:50:1 random_spheres_demo:400:225:256:50:4 \
   stress_spheres@1000:1920:1080:4:50:1 stress_spheres@1000:1920:1080:4:50:4 >> gpurun_out/r2_sweep_3.log 2>&1
cat gpurun_out/r2_sweep_3.log
python bench.py > gpurun_out/r2_bench_3.log 2> gpurun_out/r2_bench_3.err
cat gpurun_out/r2_bench_3.log | cut -c1-4000
python bench.py --impl reference --steps 2 > gpurun_out/r2_bench_3_ref.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/r2_ncu_launch.log 2>&1; echo "launch list rc=$?"
python scripts/render_once.py cornell 1000 0 > gpurun_out/r2_plain_final.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_warpq -s 1 -c 1 -f -o gpurun_out/prof_r2_cornell_warpq \
    python scripts/render_once.py cornell 1000 0 > gpurun_out/r2_ncu_final.log 2>&1; echo "full rc=$?"
python scripts/render_once.py cornell_smoke 500 0 > gpurun_out/r2_plain_smoke.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_warpq -s 1 -c 1 -f -o gpurun_out/prof_r2_smoke_warpq \
    python scripts/render_once.py cornell_smoke 500 0 > gpurun_out/r2_ncu_smoke.log 2>&1; echo "full rc=$?"
cat gpurun_out/r2_plain_final.log gpurun_out/r2_plain_smoke.log
