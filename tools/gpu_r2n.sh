# round 2, call n: 8 CTAs per SM (64 registers) against the shipped 7 (72 registers)
set -x
L=phosphorus_mk2_b200/lib
timeout 600 python tools/sweep.py --workloads spheres,terrain_bounce,terrain_nee --steps 6 $L/libphos_cuda.so $L/libphos_cuda_mb8.so $L/libphos_cuda.so $L/libphos_cuda_mb8.so 2>&1 | grep -v Adding | tee gpurun_out/sweep_r2n.log
