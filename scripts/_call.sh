set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 1500 python -m pytest tests -m gpu -q -rxXs > gpurun_out/r2_pytest_gpu_5.log 2>&1
tail -6 gpurun_out/r2_pytest_gpu_5.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2_smoke.log 2>&1; tail -2 gpurun_out/r2_smoke.log
python bench.py > gpurun_out/r2_bench_5.log 2> gpurun_out/r2_bench_5.err
cat gpurun_out/r2_bench_5.log | cut -c1-1500
python scripts/_sweep.py default final_scene:800:800:64:100:0 random_spheres_demo:400:225:256:50:0 stress_spheres@1000:1920:1080:4:50:0 bowser_demo:400:225:64:50:0 bowser_demo:400:225:64:50:1 balls_demo:400:225:64:50:0 balls_demo:400:225:64:50:1 perlin_demo:400:225:64:50:0 perlin_demo:400:225:64:50:1 > gpurun_out/r2_sweep_8.log 2>&1; cat gpurun_out/r2_sweep_8.log
for job in "final_scene 16 5" "random_spheres 256 5" "stress_1m 2 5 1920"; do
  set -- $job
  python scripts/render_once.py $job > gpurun_out/r2_plain_sq_$1.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:k_stepq -s 1 -c 1 -f -o gpurun_out/prof_r2_sq_$1 \
      python scripts/render_once.py $job > gpurun_out/r2_ncu_sq_$1.log 2>&1; echo "full rc=$?"
  cat gpurun_out/r2_plain_sq_$1.log
done
