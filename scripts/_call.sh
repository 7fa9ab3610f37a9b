# sanity of the from-scratch rebuild (make clean; make): smoke and a short bench line
set -x
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-other-configs 2> gpurun_out/bench_short.err | cut -c1-700
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "variants_equal_megakernel or hit_parity_fast_math or render_build_hits_equal" 2>&1 | tail -3
