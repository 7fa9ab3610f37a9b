# reduced validation of the build with Philox's products as 64-bit multiplies (bit-identical arithmetic): smoke, the Philox KATs,
# variant equality, fast-vs-strict hits, determinism, a short bench line, then the furnace tests while time remains
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "philox or variants_equal_megakernel or hit_parity_fast_math or render_build_hits_equal or deterministic or spp_slices" 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-other-configs 2> gpurun_out/bench_short.err | tee gpurun_out/bench_short.json | cut -c1-420
timeout 35 python -m pytest tests/test_zz_furnace_gpu.py -m gpu -q -x 2>&1 | tail -2
