set -x
for i in 1 2; do timeout 300 python scripts/_sweep.py inlmed2 cornell_box:600:600:1000:100:0 cornell_smoke:600:600:500:100:0 perlin_demo:600:600:64:50:0; done
python -m pytest tests -m gpu -q -x -k "smoke or furnace or variants_equal or config_sized" 2>&1 | tail -3
