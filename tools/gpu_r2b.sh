# round 2, call b: guided claims + tail work sharing: correctness (GPU trace tests, full parity) and the variants against each other
set -x
L=phosphorus_mk2_b200/lib
( time python -m pytest tests/test_gpu_trace.py tests/test_gpu_render.py -m gpu -x -q ) 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_r2b.log
python tools/sweep.py --workloads spheres,terrain_bounce,terrain_nee --steps 6 $L/libphos_cuda_base.so $L/libphos_cuda_div.so $L/libphos_cuda_share.so $L/libphos_cuda.so $L/libphos_cuda_d4m8.so $L/libphos_cuda_d1.so $L/libphos_cuda_base.so 2>&1 | grep -v Adding | tee gpurun_out/sweep_r2b.log
rm -f /tmp/probe_*.bin
python tools/sweep.py --workloads terrain_bounce --steps 4 $L/libphos_cuda_probe.so:PHOS_TAIL_PROBE_FILE=/tmp/probe_bounce.bin 2>&1 | grep -v Adding | tee gpurun_out/probe_r2b.log
python tools/tail_probe.py /tmp/probe_bounce.bin 400000 | tail -4 | tee -a gpurun_out/probe_r2b.log
(python tools/full_parity.py terrain; python tools/full_parity.py spheres) 2>&1 | grep -v "Adding material" | tee gpurun_out/full_parity_r2b.log
