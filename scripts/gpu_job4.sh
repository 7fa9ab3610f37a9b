#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu.log
for v in 1 2; do for c in "cornell 1000" "cornell_smoke 200"; do echo "variant $v: $(timeout 300 python scripts/render_once.py $c $v 2>&1 | tail -1)"; done; done 2>&1 | tee gpurun_out/configs_v5.log
