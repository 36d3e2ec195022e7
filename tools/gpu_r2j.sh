# round 2, call j: windowed shading classes with the Russian-roulette bit; window / register variants
set -x
L=phosphorus_mk2_b200/lib
( time timeout 400 python -m pytest tests/test_gpu_render.py tests/test_gpu_integration.py -m gpu -q -x --tb=short ) 2>&1 | grep -v "^[0-9]*, $\|Adding material" | tail -8 | tee gpurun_out/pytest_gpu_r2j.log
for v in "" _a0 _w2 _b3 _w2b3 _w8b3; do for w in cornell terrain_ggx; do
  PHOS_CUDA_LIB=$PWD/$L/libphos_cuda$v.so timeout 300 python bench.py --render --workload $w --spp 64 --depth 8 --steps 6 --warmup 2 2>/dev/null | tail -1 | cut -c1-120 | sed "s/^/lib=$v $w /" | tee -a gpurun_out/render_r2j.log
done; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_render_cornell_r2j.csv python bench.py --render --workload cornell --spp 64 --depth 8 --steps 1 --warmup 0 > gpurun_out/ncu_l.log 2>&1
python tools/launch_summary.py gpurun_out/launches_render_cornell_r2j.csv | tee gpurun_out/launch_summary_cornell_r2j.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:integrate_kernel -s 2 -c 1 -f -o gpurun_out/prof_integrate_r2j python bench.py --render --workload cornell --spp 64 --depth 8 --steps 1 --warmup 0 > gpurun_out/ncu_i.log 2>&1
