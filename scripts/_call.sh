set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
BVH="final_scene:800:800:64:100 random_spheres_demo:400:225:256:50 stress_spheres@1000:1920:1080:4:50"
rm -f gpurun_out/r2_sweep_4.log
# first: does the step-queue kernel render the right frames (strict build, bit-identical to the megakernel)?
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "variants_equal_megakernel and stepq" > gpurun_out/r2_pytest_stepq.log 2>&1; tail -5 gpurun_out/r2_pytest_stepq.log
J=""; for b in $BVH; do J="$J $b:1 $b:4 $b:5"; done
timeout 600 python scripts/_sweep.py default $J >> gpurun_out/r2_sweep_4.log 2>&1
for tag in sq64 sqn1 sqn3 sqsd4; do
  J=""; for b in $BVH; do J="$J $b:5"; done
  VECCHIO_GPU_LIB=build/libvk_$tag.so timeout 300 python scripts/_sweep.py $tag $J >> gpurun_out/r2_sweep_4.log 2>&1
done
for tag in bvh64 bvh48; do
  J=""; for b in $BVH; do J="$J $b:4"; done
  VECCHIO_GPU_LIB=build/libvk_$tag.so timeout 300 python scripts/_sweep.py $tag $J >> gpurun_out/r2_sweep_4.log 2>&1
done
for tag in base f6 f6k1 f8k2; do
  VECCHIO_GPU_LIB=build/libvk_$tag.so timeout 300 python scripts/_sweep.py $tag cornell_box:600:600:1000:100:4 cornell_box:600:600:125:100:4 cornell_smoke:600:600:500:100:4 >> gpurun_out/r2_sweep_4.log 2>&1
done
cat gpurun_out/r2_sweep_4.log
timeout 1500 python -m pytest tests -m gpu -q -rxXs > gpurun_out/r2_pytest_gpu_4.log 2>&1
tail -8 gpurun_out/r2_pytest_gpu_4.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2_smoke.log 2>&1; tail -2 gpurun_out/r2_smoke.log
python bench.py > gpurun_out/r2_bench_4.log 2> gpurun_out/r2_bench_4.err
cat gpurun_out/r2_bench_4.log | cut -c1-3000
