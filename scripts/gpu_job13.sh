#!/bin/bash
mkdir -p gpurun_out
for c in "final_scene 64" "random_spheres 256" "stress_1m 2"; do
  echo "dyn s4 a22: $(timeout 300 python scripts/render_once.py $c 1 2>&1 | tail -1)"
  for lib in build/variants/lib_dyn_*.so; do echo "$lib: $(VECCHIO_GPU_LIB=$PWD/$lib timeout 300 python scripts/render_once.py $c 1 2>&1 | tail -1)"; done
  echo "static:     $(VECCHIO_MEGA=static timeout 300 python scripts/render_once.py $c 1 2>&1 | tail -1)"
done 2>&1 | tee gpurun_out/configs_dyn2.log
