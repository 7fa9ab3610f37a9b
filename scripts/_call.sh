# the round's validation (new defaults: lane megakernels at 5 CTAs per SM, VKQ_REGEN_MIN 32), then sweep 24
set -x
bash scripts/gpu_round.sh
{
echo "# sweep 24: (product) = the new defaults; k64b10: lane megakernels as 64-thread CTAs, 10 per SM; base / rg32: VKQ_REGEN_MIN 16 / 32 with the OLD lane megakernels (6 CTAs), here for the step-queue scenes"
for round in 1 2; do
  timeout 120 python scripts/_sweep.py product final_scene:800:800:64:100:0 bowser_demo:600:600:64:100:0 random_spheres_demo:400:225:16:50:0 random_spheres_demo:400:225:256:50:0
  VECCHIO_GPU_LIB=build/libvk_k64b10.so timeout 120 python scripts/_sweep.py k64b10 final_scene:800:800:64:100:0 bowser_demo:600:600:64:100:0 random_spheres_demo:400:225:16:50:0
  for t in base rg32; do
    VECCHIO_GPU_LIB=build/libvk_$t.so timeout 120 python scripts/_sweep.py $t random_spheres_demo:400:225:256:50:0 stress_spheres@1000:1920:1080:4:50:0
  done
done
} > gpurun_out/r2_sweep_24.log 2>&1
grep -v "^+" gpurun_out/r2_sweep_24.log
