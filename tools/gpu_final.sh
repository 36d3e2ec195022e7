# end-of-round check on the GPU box: the GPU suite, smoke(), the default bench line, the two render records
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/pytest_gpu_final.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee gpurun_out/smoke_final.log
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 300 gpurun_out/bench_final.json
python bench.py --render --workload cornell --spp 64 --depth 8 > gpurun_out/render_final_cornell.json 2>/dev/null; tail -c 250 gpurun_out/render_final_cornell.json
python bench.py --render --workload terrain_ggx --spp 64 --depth 8 --steps 8 > gpurun_out/render_final_config4.json 2>/dev/null; tail -c 250 gpurun_out/render_final_config4.json
