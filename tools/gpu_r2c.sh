# round 2, call c: guided claims (fresh cursor) + tail work sharing in a separate loop; the new bench line
set -x
L=phosphorus_mk2_b200/lib
( time python -m pytest tests/test_gpu_trace.py -m gpu -x -q ) 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_r2c.log
python tools/sweep.py --workloads spheres,terrain_bounce,terrain_nee --steps 6 $L/libphos_cuda_base.so $L/libphos_cuda_div.so $L/libphos_cuda_share.so $L/libphos_cuda.so $L/libphos_cuda_d4.so $L/libphos_cuda_d1.so $L/libphos_cuda_d3m8.so $L/libphos_cuda_base.so 2>&1 | grep -v Adding | tee gpurun_out/sweep_r2c.log
rm -f /tmp/probe_*.bin
python tools/sweep.py --workloads terrain_bounce --steps 4 $L/libphos_cuda_probe.so:PHOS_TAIL_PROBE_FILE=/tmp/probe_bounce.bin 2>&1 | grep -v Adding | tee gpurun_out/probe_r2c.log
python tools/tail_probe.py /tmp/probe_bounce.bin 400000 | tail -4 | tee -a gpurun_out/probe_r2c.log
( time python bench.py ) > gpurun_out/bench_r2c.json 2> gpurun_out/bench_r2c.err; tail -c 1500 gpurun_out/bench_r2c.json; tail -5 gpurun_out/bench_r2c.err
(python tools/full_parity.py terrain) 2>&1 | grep -v "Adding material" | tee gpurun_out/full_parity_r2c.log
