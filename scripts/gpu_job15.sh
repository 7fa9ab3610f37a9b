#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "variants or edge or wavefront" > gpurun_out/pytest_wf.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_wf.log
for c in "final_scene 64" "random_spheres 256" "stress_1m 2"; do
  echo "wavefront: $(timeout 300 python scripts/render_once.py $c 2 2>&1 | tail -1)"
  echo "wf 1M slots: $(VECCHIO_WF_SLOTS=1048576 timeout 300 python scripts/render_once.py $c 2 2>&1 | tail -1)"
  echo "wf 2M slots: $(VECCHIO_WF_SLOTS=2097152 timeout 300 python scripts/render_once.py $c 2 2>&1 | tail -1)"
  echo "auto:      $(timeout 300 python scripts/render_once.py $c 0 2>&1 | tail -1)"
done 2>&1 | tee gpurun_out/configs_wfdyn.log
