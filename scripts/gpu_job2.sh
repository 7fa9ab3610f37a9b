#!/bin/bash
# wavefront bring-up: tests, then timing of both variants on the configs
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "wavefront or edge" > gpurun_out/pytest_wf.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_wf.log
for v in 1 2; do for c in "cornell 1000" "cornell_smoke 200" "final_scene 64" "random_spheres 16"; do echo "variant $v: $(timeout 300 python scripts/render_once.py $c $v 2>&1 | tail -1)"; done; done 2>&1 | tee gpurun_out/wf_configs.log
for n in 131072 262144 1048576 2097152; do echo "slots $n: $(VECCHIO_WF_SLOTS=$n timeout 300 python scripts/render_once.py cornell 1000 2 2>&1 | tail -1)"; done 2>&1 | tee -a gpurun_out/wf_configs.log
