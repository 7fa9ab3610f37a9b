#!/bin/bash
# usage: scripts/sweep_variants.sh <config> <spp>  -- time every library under build/variants plus the default build
for lib in "" build/variants/*.so; do
  if [ -n "$lib" ]; then export VECCHIO_GPU_LIB=$PWD/$lib; else unset VECCHIO_GPU_LIB; fi
  echo "== ${lib:-default}: $(python scripts/render_once.py $1 $2 2>&1 | tail -1)"
done
