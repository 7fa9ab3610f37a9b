set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
python -m pytest tests -m gpu -q -rxXs > gpurun_out/r2_pytest_gpu_2.log 2>&1
tail -5 gpurun_out/r2_pytest_gpu_2.log
FLAT="cornell_box:600:600:1000:100:4 cornell_smoke:600:600:500:100:4"
BVH="final_scene:800:800:64:100:4 random_spheres_demo:400:225:256:50:4 stress_spheres@1000:1920:1080:4:50:4 bowser_demo:600:337:64:50:4"
rm -f gpurun_out/r2_sweep_2.log
for tag in a2 a2fr a3fr a1 c2 c1 a2nocv; do
  VECCHIO_GPU_LIB=build/libvk_$tag.so python scripts/_sweep.py $tag $FLAT >> gpurun_out/r2_sweep_2.log 2>&1
done
for tag in a2 c2 a2i4 a2i16 a2n2 a2nocv; do
  VECCHIO_GPU_LIB=build/libvk_$tag.so python scripts/_sweep.py $tag $BVH >> gpurun_out/r2_sweep_2.log 2>&1
done
python scripts/_sweep.py default cornell_box:600:600:1000:100:3 cornell_smoke:600:600:500:100:3 cornell_box:600:600:1000:100:4 cornell_smoke:600:600:500:100:4 \
   final_scene:800:800:64:100:1 final_scene:800:800:64:100:4 random_spheres_demo:400:225:256:50:1 random_spheres_demo:400:225:256:50:4 \
   stress_spheres@1000:1920:1080:4:50:1 stress_spheres@1000:1920:1080:4:50:4 bowser_demo:600:337:64:50:1 bowser_demo:600:337:64:50:4 >> gpurun_out/r2_sweep_2.log 2>&1
VECCHIO_L2_PERSIST=0 python scripts/_sweep.py nopersist stress_spheres@1000:1920:1080:4:50:1 stress_spheres@1000:1920:1080:4:50:4 stress_spheres@1000:3840:2160:2:50:4 stress_spheres@1000:3840:2160:2:50:1 >> gpurun_out/r2_sweep_2.log 2>&1
python scripts/_sweep.py persist stress_spheres@1000:3840:2160:2:50:4 stress_spheres@1000:3840:2160:2:50:1 >> gpurun_out/r2_sweep_2.log 2>&1
cat gpurun_out/r2_sweep_2.log
timeout 300 compute-sanitizer --tool racecheck python scripts/render_once.py cornell 2 4 96 > gpurun_out/r2_racecheck_flat.log 2>&1; tail -4 gpurun_out/r2_racecheck_flat.log
timeout 300 compute-sanitizer --tool racecheck python scripts/render_once.py final_scene 1 4 64 > gpurun_out/r2_racecheck_bvh.log 2>&1; tail -4 gpurun_out/r2_racecheck_bvh.log
timeout 300 compute-sanitizer --tool memcheck python scripts/render_once.py cornell_smoke 2 4 96 > gpurun_out/r2_memcheck_flat.log 2>&1; tail -4 gpurun_out/r2_memcheck_flat.log
python scripts/render_once.py cornell 64 4 > gpurun_out/r2_plain_wq.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_warpq -s 1 -c 1 -f -o gpurun_out/prof_r2_warpq_v1 \
    python scripts/render_once.py cornell 64 4 > gpurun_out/r2_ncu_wq.log 2>&1; echo "full rc=$?"
python scripts/render_once.py final_scene 16 4 > gpurun_out/r2_plain_wq_fs.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_warpq -s 1 -c 1 -f -o gpurun_out/prof_r2_warpq_fs_v1 \
    python scripts/render_once.py final_scene 16 4 > gpurun_out/r2_ncu_wq_fs.log 2>&1; echo "full rc=$?"
python scripts/render_once.py stress_1m 2 4 1920 > gpurun_out/r2_plain_wq_st.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_warpq -s 1 -c 1 -f -o gpurun_out/prof_r2_warpq_st_v1 \
    python scripts/render_once.py stress_1m 2 4 1920 > gpurun_out/r2_ncu_wq_st.log 2>&1; echo "full rc=$?"
python bench.py > gpurun_out/r2_bench_2.log 2> gpurun_out/r2_bench_2.err
cat gpurun_out/r2_bench_2.log | cut -c1-3000
