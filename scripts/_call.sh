set -x
for i in 1 2; do timeout 300 python scripts/_sweep.py nospec final_scene:800:800:64:100:0 random_spheres_demo:400:225:256:50:0 balls_demo:600:600:64:50:0 bowser_demo:600:600:64:50:0 perlin_demo:600:600:64:50:0; done
