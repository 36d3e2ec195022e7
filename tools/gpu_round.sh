# GPU-box round: tests, bench (+ reference arm), launch list, ncu captures of the traversal kernel, full parity,
# render benches (suffix = $1)
S=${1:-r1m}
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu_$S.log
python bench.py > gpurun_out/bench_$S.json 2> gpurun_out/bench_$S.err; tail -c 3000 gpurun_out/bench_$S.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$S.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 5 -c 1 -f -o gpurun_out/prof_trace_$S python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 4 -c 1 -f -o gpurun_out/prof_terrain_bounce_$S python tools/sweep.py --workloads terrain_bounce --steps 2 phosphorus_mk2_b200/lib/libphos_cuda.so > gpurun_out/ncu3.log 2>&1
ls -la gpurun_out/*_$S.ncu-rep
python bench.py --impl reference > gpurun_out/bench_${S}_reference.json 2>/dev/null; tail -c 600 gpurun_out/bench_${S}_reference.json
python bench.py --render --workload cornell --spp 64 --depth 8 > gpurun_out/render_${S}_cornell.json 2>/dev/null; tail -c 700 gpurun_out/render_${S}_cornell.json
python bench.py --render --workload terrain_ggx --spp 64 --depth 8 > gpurun_out/render_${S}_config4.json 2>/dev/null; tail -c 700 gpurun_out/render_${S}_config4.json
(python tools/full_parity.py terrain; python tools/full_parity.py spheres) 2>&1 | grep -v "Adding material" | tee gpurun_out/full_parity_$S.log
