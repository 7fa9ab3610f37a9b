set -x
timeout 300 python scripts/_sweep.py hitc cornell_box:600:600:1000:100:0 cornell_smoke:600:600:500:100:0 perlin_demo:600:600:64:50:0 balls_demo:600:600:64:50:0
python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
