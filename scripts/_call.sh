set -x
BVH="final_scene:800:800:64:100 random_spheres_demo:400:225:256:50 stress_spheres@1000:1920:1080:4:50"
rm -f gpurun_out/r2_sweep_5.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_zz_warpq_selfcheck_gpu.py -m gpu -q -x -k "stepq or selfcheck or overflow" > gpurun_out/r2_pytest_stepq.log 2>&1; tail -5 gpurun_out/r2_pytest_stepq.log
J=""; for b in $BVH; do J="$J $b:1 $b:5"; done
timeout 600 python scripts/_sweep.py default $J >> gpurun_out/r2_sweep_5.log 2>&1
for tag in nop nos nom mw sd4 n128 ns4; do
  J=""; for b in $BVH; do J="$J $b:5"; done
  VECCHIO_GPU_LIB=build/libvk_$tag.so timeout 300 python scripts/_sweep.py $tag $J >> gpurun_out/r2_sweep_5.log 2>&1
done
cat gpurun_out/r2_sweep_5.log
