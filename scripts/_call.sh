set -x
for i in 1 2; do VECCHIO_GPU_LIB=build/libvk_n176c.so timeout 300 python scripts/_sweep.py n176c cornell_box:600:600:1000:100:0; done
