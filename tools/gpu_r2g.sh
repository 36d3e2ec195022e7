# round 2, call g: window-local shading classes (+ window size, register cap of integrate), tests
set -x
L=phosphorus_mk2_b200/lib
( time timeout 300 python -m pytest tests/test_gpu_integration.py tests/test_gpu_render.py -m gpu -q --tb=short ) 2>&1 | grep -v "^[0-9]*, $\|Adding material" | tail -30 | tee gpurun_out/pytest_gpu_r2g.log
for v in "" _w8 _w2 _b3 _b2 _b5; do for w in cornell terrain_ggx; do
  PHOS_CUDA_LIB=$PWD/$L/libphos_cuda$v.so timeout 300 python bench.py --render --workload $w --spp 64 --depth 8 --steps 6 --warmup 2 2>/dev/null | tail -1 | cut -c1-120 | sed "s/^/lib=$v $w /" | tee -a gpurun_out/render_bin_r2g.log
done; done
for w in cornell terrain_ggx; do
  PHOS_SHADE_BIN=0 timeout 300 python bench.py --render --workload $w --spp 64 --depth 8 --steps 6 --warmup 2 2>/dev/null | tail -1 | cut -c1-120 | sed "s/^/bin=0 $w /" | tee -a gpurun_out/render_bin_r2g.log
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_render_cornell_r2g.csv python bench.py --render --workload cornell --spp 64 --depth 8 --steps 1 --warmup 0 > gpurun_out/ncu_l.log 2>&1
python tools/launch_summary.py gpurun_out/launches_render_cornell_r2g.csv | tee gpurun_out/launch_summary_cornell_r2g.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:integrate_kernel -s 2 -c 1 -f -o gpurun_out/prof_integrate_r2g python bench.py --render --workload cornell --spp 64 --depth 8 --steps 1 --warmup 0 > gpurun_out/ncu_i.log 2>&1
