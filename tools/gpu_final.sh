# end-of-round check on the GPU box: the GPU suite, smoke(), the default bench line and the reference arm
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/pytest_gpu_final.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee gpurun_out/smoke_final.log
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 1200 gpurun_out/bench_final.json
python bench.py --impl reference --steps 5 --warmup 1 2>/dev/null | tail -c 400
