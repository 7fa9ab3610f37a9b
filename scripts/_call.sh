set -x
timeout 300 python scripts/_sweep.py l0 final_scene:800:800:64:100:0 random_spheres_demo:400:225:256:50:0 balls_demo:600:600:64:50:0 api_surface_demo:600:600:64:50:0 stress_spheres@1000:1920:1080:4:50:0 cornell_box:600:600:1000:100:0
python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
