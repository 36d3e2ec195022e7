S=${1:-r1n}
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/pytest_gpu_$S.log
python bench.py > gpurun_out/bench_$S.json 2> gpurun_out/bench_$S.err; tail -c 2500 gpurun_out/bench_$S.json
python bench.py --impl reference > gpurun_out/bench_${S}_reference.json 2>/dev/null; tail -c 300 gpurun_out/bench_${S}_reference.json
python bench.py --render --workload cornell --spp 64 --depth 8 > gpurun_out/render_${S}_cornell.json 2>/dev/null; tail -c 300 gpurun_out/render_${S}_cornell.json
python bench.py --render --workload terrain_ggx --spp 64 --depth 8 > gpurun_out/render_${S}_config4.json 2>/dev/null; tail -c 300 gpurun_out/render_${S}_config4.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_render_config4_$S.csv python bench.py --render --workload terrain_ggx --spp 64 --depth 8 --steps 1 --warmup 0 > gpurun_out/ncu4.log 2>&1
python tools/launch_summary.py gpurun_out/launches_render_config4_$S.csv
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_render_cornell_$S.csv python bench.py --render --workload cornell --spp 64 --depth 8 --steps 1 --warmup 0 > gpurun_out/ncu5.log 2>&1
python tools/launch_summary.py gpurun_out/launches_render_cornell_$S.csv
