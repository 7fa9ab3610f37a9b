#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_gpu.log
for c in "final_scene 64" "random_spheres 16" "random_spheres 256" "stress_1m 2"; do
  echo "dyn:    $(timeout 300 python scripts/render_once.py $c 1 2>&1 | tail -1)"
  echo "static: $(VECCHIO_MEGA=static timeout 300 python scripts/render_once.py $c 1 2>&1 | tail -1)"
done 2>&1 | tee gpurun_out/configs_dyn.log
