set -x
J="cornell_box:600:600:1000:100:0 cornell_smoke:600:600:500:100:0"
rm -f gpurun_out/r2_sweep_11.log
python scripts/_sweep.py slabbox $J >> gpurun_out/r2_sweep_11.log 2>&1
for tag in n120b4 n120b6; do
  VECCHIO_GPU_LIB=build/libvk_$tag.so timeout 300 python scripts/_sweep.py $tag cornell_box:600:600:1000:100:0 >> gpurun_out/r2_sweep_11.log 2>&1
done
cat gpurun_out/r2_sweep_11.log
python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
