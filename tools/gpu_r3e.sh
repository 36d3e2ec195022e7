# round 2, call 3e: the eager-commit kernel (default build) through the GPU trace tests; chunk sizes 16 / 24 / 32 rays and refill thresholds on it
set -x
L=phosphorus_mk2_b200/lib
( timeout 300 python -m pytest tests/test_gpu_trace.py -m gpu -q -x --tb=short ) 2>&1 | tail -3 | tee gpurun_out/pytest_gpu_r3e.log
timeout 900 python tools/sweep.py --workloads spheres,terrain_bounce,terrain_nee --steps 8 $L/libphos_cuda.so $L/libphos_cuda_c16.so $L/libphos_cuda_c24.so $L/libphos_cuda_c16r8.so $L/libphos_cuda_r8.so $L/libphos_cuda_r10.so $L/libphos_cuda.so $L/libphos_cuda_c16.so $L/libphos_cuda_r8.so 2>&1 | grep -v Adding | tee gpurun_out/sweep_r3e.log
