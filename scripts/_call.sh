set -x
J="final_scene:800:800:64:100:0 random_spheres_demo:400:225:256:50:0 perlin_demo:600:600:64:50:0 balls_demo:600:600:64:50:0 stress_spheres@1000:1920:1080:4:50:0 bowser_demo:600:600:64:50:0 api_surface_demo:600:600:64:50:0"
for i in 1 2; do
timeout 300 python scripts/_sweep.py light0 $J
VECCHIO_GPU_LIB=build/libvk_nolight0.so timeout 300 python scripts/_sweep.py nolight0 $J
done
