# round 2, call e: sharing variants; the failing integration test in full
set -x
L=phosphorus_mk2_b200/lib
timeout 600 python tools/sweep.py --workloads spheres,terrain_bounce,terrain_nee --steps 6 $L/libphos_cuda_base.so $L/libphos_cuda.so $L/libphos_cuda_s1.so $L/libphos_cuda_s2.so $L/libphos_cuda_s3.so $L/libphos_cuda_s4.so $L/libphos_cuda_s5.so $L/libphos_cuda_base.so 2>&1 | grep -v Adding | tee gpurun_out/sweep_r2e.log
( time timeout 300 python -m pytest tests/test_gpu_integration.py tests/test_gpu_render.py tests/test_gpu_multi.py -m gpu -q ) 2>&1 | grep -v "^[0-9]*, $" | tail -60 | tee gpurun_out/pytest_gpu_r2e.log
