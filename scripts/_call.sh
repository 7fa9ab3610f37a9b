set -x
rm -f gpurun_out/r2_sweep_14.log
timeout 300 python scripts/_sweep.py span cornell_box:600:600:1000:100:0 cornell_smoke:600:600:500:100:0 final_scene:800:800:64:100:0 final_scene:800:800:64:100:4 final_scene:800:800:64:100:5 >> gpurun_out/r2_sweep_14.log 2>&1
cat gpurun_out/r2_sweep_14.log
python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
