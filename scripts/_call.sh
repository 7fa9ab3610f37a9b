set -x
BVH="final_scene:800:800:64:100 random_spheres_demo:400:225:256:50 stress_spheres@1000:1920:1080:4:50 random_spheres_demo:400:225:16:50"
rm -f gpurun_out/r2_sweep_10.log
J=""; for b in $BVH; do J="$J $b:5"; done
for tag in is1n6 is1n8 is1n12 is0n6 is1n6l3 n6w5; do
  VECCHIO_GPU_LIB=build/libvk_$tag.so timeout 300 python scripts/_sweep.py $tag $J >> gpurun_out/r2_sweep_10.log 2>&1
done
cat gpurun_out/r2_sweep_10.log
