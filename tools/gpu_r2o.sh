# round 2, call o: cheaper drain-loop protocol (pre-check, sharing every n-th iteration); "never" = two-loop structure without sharing
set -x
L=phosphorus_mk2_b200/lib
( timeout 300 python -m pytest tests/test_gpu_trace.py -m gpu -q -x --tb=short ) 2>&1 | tail -3 | tee gpurun_out/pytest_gpu_r2o.log
timeout 700 python tools/sweep.py --workloads spheres,terrain_bounce,terrain_nee --steps 6 $L/libphos_cuda_base.so $L/libphos_cuda.so $L/libphos_cuda_e2.so $L/libphos_cuda_e4.so $L/libphos_cuda_never.so $L/libphos_cuda_base.so $L/libphos_cuda.so 2>&1 | grep -v Adding | tee gpurun_out/sweep_r2o.log
