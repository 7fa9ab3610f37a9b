set -x
for i in 1 2; do
timeout 300 python scripts/_sweep.py peel cornell_box:600:600:1000:100:0 cornell_smoke:600:600:500:100:0 perlin_demo:600:600:64:50:0 balls_demo:600:600:64:50:0
VECCHIO_GPU_LIB=build/libvk_nopeel.so timeout 300 python scripts/_sweep.py nopeel cornell_box:600:600:1000:100:0 cornell_smoke:600:600:500:100:0 perlin_demo:600:600:64:50:0 balls_demo:600:600:64:50:0
done
python -m pytest tests -m gpu -q -x -k "variants_equal or hybrid or fast_math or render_build" 2>&1 | tail -3
