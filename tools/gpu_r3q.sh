# round 2, call 3q: deferred leaf sets (df: vote bias 3, dfb2: bias 2) against the shipped kernel; GPU trace tests on the df build
set -x
L=$PWD/phosphorus_mk2_b200/lib
( PHOS_CUDA_LIB=$L/libphos_cuda_df.so timeout 300 python -m pytest tests/test_gpu_trace.py -m gpu -q -x --tb=short ) 2>&1 | tail -3 | tee gpurun_out/pytest_gpu_r3q.log
timeout 900 python tools/sweep.py --workloads spheres,terrain_bounce,terrain_nee --steps 8 $L/libphos_cuda.so $L/libphos_cuda_df.so $L/libphos_cuda_dfb2.so $L/libphos_cuda.so $L/libphos_cuda_df.so $L/libphos_cuda_dfb2.so 2>&1 | grep -v Adding | tee gpurun_out/sweep_r3q.log
