set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu_r1j.log
python bench.py > gpurun_out/bench_r1j.json 2> gpurun_out/bench_r1j.err; tail -c 3000 gpurun_out/bench_r1j.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1j.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 5 -c 1 -f -o gpurun_out/prof_trace_r1j python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 4 -c 1 -f -o gpurun_out/prof_terrain_bounce_r1j python tools/sweep.py --workloads terrain_bounce --steps 2 phosphorus_mk2_b200/lib/libphos_cuda.so > gpurun_out/ncu3.log 2>&1
ls -la gpurun_out/*.ncu-rep
python bench.py --impl reference > gpurun_out/bench_r1j_reference.json 2>/dev/null; tail -c 600 gpurun_out/bench_r1j_reference.json
python bench.py --render --workload cornell --spp 64 --depth 8 > gpurun_out/render_r1j_cornell.json 2>/dev/null; tail -c 700 gpurun_out/render_r1j_cornell.json
python bench.py --render --workload terrain_ggx --spp 64 --depth 8 > gpurun_out/render_r1j_config4.json 2>/dev/null; tail -c 700 gpurun_out/render_r1j_config4.json
