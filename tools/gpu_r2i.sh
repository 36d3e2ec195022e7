# round 2, call i: three-phase integrate (requests sorted by shading class in shared memory); tests
set -x
L=phosphorus_mk2_b200/lib
( time timeout 400 python -m pytest tests/test_gpu_render.py tests/test_gpu_integration.py -m gpu -q -x --tb=short ) 2>&1 | grep -v "^[0-9]*, $\|Adding material" | tail -30 | tee gpurun_out/pytest_gpu_r2i.log
for v in "" _b3 _b2 _b5; do for w in cornell terrain_ggx; do
  PHOS_CUDA_LIB=$PWD/$L/libphos_cuda$v.so timeout 300 python bench.py --render --workload $w --spp 64 --depth 8 --steps 6 --warmup 2 2>/dev/null | tail -1 | cut -c1-120 | sed "s/^/lib=$v $w /" | tee -a gpurun_out/render_r2i.log
done; done
for w in cornell terrain_ggx; do
  PHOS_SHADE_BIN=0 timeout 300 python bench.py --render --workload $w --spp 64 --depth 8 --steps 6 --warmup 2 2>/dev/null | tail -1 | cut -c1-120 | sed "s/^/bin=0 $w /" | tee -a gpurun_out/render_r2i.log
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_render_cornell_r2i.csv python bench.py --render --workload cornell --spp 64 --depth 8 --steps 1 --warmup 0 > gpurun_out/ncu_l.log 2>&1
python tools/launch_summary.py gpurun_out/launches_render_cornell_r2i.csv | tee gpurun_out/launch_summary_cornell_r2i.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:integrate_kernel -s 2 -c 1 -f -o gpurun_out/prof_integrate_r2i python bench.py --render --workload cornell --spp 64 --depth 8 --steps 1 --warmup 0 > gpurun_out/ncu_i.log 2>&1
