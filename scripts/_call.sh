set -x
for job in "final_scene 16 5" "random_spheres 256 5" "stress_1m 2 5 1920"; do
  set -- $job
  python scripts/render_once.py $job > gpurun_out/r2_plain_sq_$1.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:k_stepq -s 1 -c 1 -f -o gpurun_out/prof_r2_sq_$1 \
      python scripts/render_once.py $job > gpurun_out/r2_ncu_sq_$1.log 2>&1; echo "full rc=$?"
  cat gpurun_out/r2_plain_sq_$1.log
done
