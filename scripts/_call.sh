set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
python -m pytest tests -m gpu -q -rxXs > gpurun_out/r2_pytest_gpu_4.log 2>&1
tail -5 gpurun_out/r2_pytest_gpu_4.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2_smoke.log 2>&1; tail -2 gpurun_out/r2_smoke.log
BVH="final_scene:800:800:64:100:4 random_spheres_demo:400:225:256:50:4 stress_spheres@1000:1920:1080:4:50:4"
FLAT="cornell_box:600:600:1000:100:4 cornell_smoke:600:600:500:100:4"
rm -f gpurun_out/r2_sweep_4.log
for tag in base bvh96 bvh64 bvh64m8 bvh48; do
  VECCHIO_GPU_LIB=build/libvk_$tag.so python scripts/_sweep.py $tag $FLAT $BVH >> gpurun_out/r2_sweep_4.log 2>&1
done
python scripts/_sweep.py default final_scene:800:800:64:100:1 random_spheres_demo:400:225:256:50:1 stress_spheres@1000:1920:1080:4:50:1 >> gpurun_out/r2_sweep_4.log 2>&1
cat gpurun_out/r2_sweep_4.log
python bench.py > gpurun_out/r2_bench_4.log 2> gpurun_out/r2_bench_4.err
cat gpurun_out/r2_bench_4.log | cut -c1-4000
