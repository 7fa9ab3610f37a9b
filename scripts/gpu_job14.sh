#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "variants or edge" > gpurun_out/pytest_staged.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_staged.log
for c in "final_scene 64" "random_spheres 256" "random_spheres 16" "stress_1m 2" "cornell_smoke 200"; do
  echo "staged: $(timeout 300 python scripts/render_once.py $c 3 2>&1 | tail -1)"
done 2>&1 | tee gpurun_out/configs_staged_bvh.log
