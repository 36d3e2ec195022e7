# round 2, call 3m: branch-free hit acceptance (default build) against the kernel before it (base)
set -x
L=phosphorus_mk2_b200/lib
( timeout 300 python -m pytest tests/test_gpu_trace.py -m gpu -q -x --tb=short ) 2>&1 | tail -3 | tee gpurun_out/pytest_gpu_r3m.log
timeout 900 python tools/sweep.py --workloads spheres,terrain_bounce,terrain_nee --steps 8 $L/libphos_cuda_base.so $L/libphos_cuda.so $L/libphos_cuda_base.so $L/libphos_cuda.so 2>&1 | grep -v Adding | tee gpurun_out/sweep_r3m.log
