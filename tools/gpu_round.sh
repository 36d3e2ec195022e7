set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu_r1d.log
python bench.py > gpurun_out/bench_r1d.json 2> gpurun_out/bench_r1d.err; tail -c 3000 gpurun_out/bench_r1d.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1d.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 5 -c 1 -f -o gpurun_out/prof_trace_r1d python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 4 -c 1 -f -o gpurun_out/prof_terrain_bounce_r1d python tools/sweep.py --workloads terrain_bounce --steps 2 phosphorus_mk2_b200/lib/libphos_cuda.so > gpurun_out/ncu3.log 2>&1
ls -la gpurun_out/*.ncu-rep
