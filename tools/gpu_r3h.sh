# round 2, call 3h: lazy staging near the end of the stream with ONE claim site in the loop (lz1 / lz2 / lz4 = last 1 / 2 / 4 chunks per warp);
# the new multi-frame drop-in test
set -x
L=phosphorus_mk2_b200/lib
( timeout 300 python -m pytest tests/test_gpu_integration.py -m gpu -q -x --tb=short ) 2>&1 | tail -3 | tee gpurun_out/pytest_gpu_r3h.log
timeout 900 python tools/sweep.py --workloads spheres,terrain_bounce,terrain_nee --steps 8 $L/libphos_cuda.so $L/libphos_cuda_lz1.so $L/libphos_cuda_lz2.so $L/libphos_cuda_lz4.so $L/libphos_cuda.so $L/libphos_cuda_lz2.so 2>&1 | grep -v Adding | tee gpurun_out/sweep_r3h.log
