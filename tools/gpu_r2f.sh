# round 2, call f: sharing threshold; shading-class binning on / off; tests; ncu of the binned integrate kernel
set -x
L=phosphorus_mk2_b200/lib
( time timeout 300 python -m pytest tests/test_gpu_integration.py tests/test_gpu_render.py tests/test_gpu_multi.py -m gpu -q --tb=short ) 2>&1 | grep -v "^[0-9]*, $\|Adding material" | tail -40 | tee gpurun_out/pytest_gpu_r2f.log
timeout 600 python tools/sweep.py --workloads spheres,terrain_bounce,terrain_nee --steps 6 $L/libphos_cuda_base.so $L/libphos_cuda.so $L/libphos_cuda_i8.so $L/libphos_cuda_i16.so $L/libphos_cuda_i20.so $L/libphos_cuda_i24.so $L/libphos_cuda_base.so 2>&1 | grep -v Adding | tee gpurun_out/sweep_r2f.log
for b in 0 1; do for w in cornell terrain_ggx; do
  PHOS_SHADE_BIN=$b timeout 300 python bench.py --render --workload $w --spp 64 --depth 8 --steps 6 --warmup 2 2>/dev/null | tail -1 | cut -c1-260 | sed "s/^/bin=$b $w /" | tee -a gpurun_out/render_bin_r2f.log
done; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_render_cornell_r2f.csv python bench.py --render --workload cornell --spp 64 --depth 8 --steps 1 --warmup 0 > gpurun_out/ncu_l.log 2>&1
python tools/launch_summary.py gpurun_out/launches_render_cornell_r2f.csv | tee gpurun_out/launch_summary_cornell_r2f.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:integrate_kernel -s 0 -c 1 -f -o gpurun_out/prof_integrate_r2f python bench.py --render --workload cornell --spp 64 --depth 8 --steps 1 --warmup 0 > gpurun_out/ncu_i.log 2>&1
ls -la gpurun_out/*r2f*
