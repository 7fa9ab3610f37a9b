#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "variants or edge" > gpurun_out/pytest_staged.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/pytest_staged.log
for v in 1 3; do for c in "cornell 1000" "cornell_smoke 200" "final_scene 64" "random_spheres 16"; do echo "variant $v: $(timeout 300 python scripts/render_once.py $c $v 2>&1 | tail -1)"; done; done 2>&1 | tee gpurun_out/configs_v6.log
