# round 2, call 3o: validation of the final state (eager commit, batched integrate loads, 128 Ki pipeline chunks) — GPU suite, smoke(), default bench line + reference arm, full-stream parity,
# ncu captures of the shipped kernels, launch lists
set -x
( time timeout 400 python -m pytest tests -m gpu -q -x --tb=short ) 2>&1 | grep -v "^[0-9]*, $\|Adding material" | tail -6 | tee gpurun_out/pytest_gpu_r3o.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('__SMOKE_OK__')" 2>&1 | grep -v "Adding material" | tail -3 | tee gpurun_out/smoke_r3o.log
( time timeout 900 python bench.py ) > gpurun_out/bench_r3o.json 2> gpurun_out/bench_r3o.err; tail -c 400 gpurun_out/bench_r3o.json; tail -4 gpurun_out/bench_r3o.err
( time timeout 600 python bench.py --impl reference ) > gpurun_out/bench_r3o_reference.json 2>/dev/null; tail -c 300 gpurun_out/bench_r3o_reference.json
(timeout 300 python tools/full_parity.py terrain; timeout 300 python tools/full_parity.py spheres) 2>&1 | grep -v "Adding material" | tee gpurun_out/full_parity_r3o.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_r3o.csv python bench.py --only headline --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 5 -c 1 -f -o gpurun_out/prof_trace_spheres_r3o python bench.py --only headline --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 4 -c 1 -f -o gpurun_out/prof_trace_bounce_r3o python tools/sweep.py --workloads terrain_bounce --steps 2 phosphorus_mk2_b200/lib/libphos_cuda.so > gpurun_out/ncu3.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 4 -c 1 -f -o gpurun_out/prof_trace_shadow_r3o python tools/sweep.py --workloads terrain_nee --steps 2 phosphorus_mk2_b200/lib/libphos_cuda.so > gpurun_out/ncu4.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_render_config4_r3o.csv python bench.py --render --workload terrain_ggx --spp 64 --depth 8 --steps 1 --warmup 0 > gpurun_out/ncu5.log 2>&1
python tools/launch_summary.py gpurun_out/launches_render_config4_r3o.csv | tee gpurun_out/launch_summary_config4_r3o.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_render_cornell_r3o.csv python bench.py --render --workload cornell --spp 64 --depth 8 --steps 1 --warmup 0 > gpurun_out/ncu6.log 2>&1
python tools/launch_summary.py gpurun_out/launches_render_cornell_r3o.csv | tee gpurun_out/launch_summary_cornell_r3o.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:integrate_kernel -s 2 -c 1 -f -o gpurun_out/prof_integrate_r3o python bench.py --render --workload cornell --spp 64 --depth 8 --steps 1 --warmup 0 > gpurun_out/ncu7.log 2>&1
ls -la gpurun_out/*r3o* | wc -l
timeout 300 ncu --set full --clock-control none --import-source on -k regex:shade_nee_kernel -s 2 -c 1 -f -o gpurun_out/prof_shade_nee_r3o python bench.py --render --workload cornell --spp 64 --depth 8 --steps 1 --warmup 0 > gpurun_out/ncu8.log 2>&1
ls -la gpurun_out/*r3o* | wc -l
