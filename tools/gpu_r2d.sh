# round 2, call d: guided claims (fresh cursor) + tail work sharing in a separate loop (every step under its own timeout)
set -x
L=phosphorus_mk2_b200/lib
( time timeout 300 python -m pytest tests/test_gpu_trace.py -m gpu -x -q ) 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_r2d.log
grep -q passed gpurun_out/pytest_gpu_r2d.log || exit 1
timeout 600 python tools/sweep.py --workloads spheres,terrain_bounce,terrain_nee --steps 6 $L/libphos_cuda_base.so $L/libphos_cuda_div.so $L/libphos_cuda_share.so $L/libphos_cuda.so $L/libphos_cuda_d4.so $L/libphos_cuda_d1.so $L/libphos_cuda_d3m8.so $L/libphos_cuda_base.so 2>&1 | grep -v Adding | tee gpurun_out/sweep_r2d.log
rm -f /tmp/probe_*.bin
timeout 200 python tools/sweep.py --workloads terrain_bounce --steps 4 $L/libphos_cuda_probe.so:PHOS_TAIL_PROBE_FILE=/tmp/probe_bounce.bin 2>&1 | grep -v Adding | tee gpurun_out/probe_r2d.log
python tools/tail_probe.py /tmp/probe_bounce.bin 400000 | tail -4 | tee -a gpurun_out/probe_r2d.log
( time timeout 300 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_all_r2d.log
( time timeout 900 python bench.py ) > gpurun_out/bench_r2d.json 2> gpurun_out/bench_r2d.err; tail -c 1500 gpurun_out/bench_r2d.json; tail -5 gpurun_out/bench_r2d.err
(timeout 300 python tools/full_parity.py terrain) 2>&1 | grep -v "Adding material" | tee gpurun_out/full_parity_r2d.log
