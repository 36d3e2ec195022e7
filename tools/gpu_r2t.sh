# round 2, call t: L1 prefetch of the first triangle of the nearest hit leaf / of the next node at the end of a node step
set -x
L=phosphorus_mk2_b200/lib
timeout 700 python tools/sweep.py --workloads spheres,terrain_bounce,terrain_nee --steps 8 $L/libphos_cuda.so $L/libphos_cuda_pt.so $L/libphos_cuda_pn.so $L/libphos_cuda_ptn.so $L/libphos_cuda.so $L/libphos_cuda_pt.so 2>&1 | grep -v Adding | tee gpurun_out/sweep_r2t.log
