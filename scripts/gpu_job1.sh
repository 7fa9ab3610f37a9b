#!/bin/bash
# round-1 GPU job: parity suite, bench (both arms), per-config timings, ncu launch list + full capture
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpu.txt; echo "nproc $(nproc)" >> gpurun_out/gpu.txt
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 3 > gpurun_out/bench_ref.log 2>&1; echo "ref rc=$?"
for c in "random_spheres 16" "cornell 1000" "cornell_smoke 200" "final_scene 64"; do python scripts/render_once.py $c 2>&1 | tail -1; done > gpurun_out/configs.log
cat gpurun_out/configs.log
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "launch list rc=$?"
python scripts/render_once.py cornell 200 > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_megakernel -s 1 -c 1 -f -o gpurun_out/prof_r1_cornell \
    python scripts/render_once.py cornell 200 > gpurun_out/ncu_full.log 2>&1; echo "full rc=$?"
