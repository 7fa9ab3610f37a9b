#!/bin/bash
# The round's GPU validation in one gpurun call: parity suite, smoke, both bench arms.
#   gpurun --timeout 1500 -- 'bash scripts/gpu_round.sh'
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpu.txt; echo "nproc $(nproc)" >> gpurun_out/gpu.txt
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
python bench.py --impl reference --steps 2 --warmup 3 > gpurun_out/bench_ref.log 2>&1; echo "reference arm rc=$?"
python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 400 gpurun_out/bench.log
