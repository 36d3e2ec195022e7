# round 2, call z: octant flag of the work sharing in shared memory (no spill) against the register version (one spill slot) and r01's loop
set -x
L=phosphorus_mk2_b200/lib
( timeout 300 python -m pytest tests/test_gpu_trace.py -m gpu -q -x --tb=short ) 2>&1 | tail -3 | tee gpurun_out/pytest_gpu_r2z.log
timeout 800 python tools/sweep.py --workloads spheres,terrain_bounce,terrain_nee --steps 8 $L/libphos_cuda_base.so $L/libphos_cuda.so $L/libphos_cuda_spill.so $L/libphos_cuda_base.so $L/libphos_cuda.so $L/libphos_cuda_spill.so 2>&1 | grep -v Adding | tee gpurun_out/sweep_r2z.log
