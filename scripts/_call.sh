set -x
for i in 1 2; do
VECCHIO_GPU_LIB=build/libvk_nospan.so timeout 300 python scripts/_sweep.py nospan final_scene:800:800:64:100:0 bowser_demo:600:600:64:50:0
timeout 300 python scripts/_sweep.py span final_scene:800:800:64:100:0 bowser_demo:600:600:64:50:0
done
