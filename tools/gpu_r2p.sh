# round 2, call p: a lane gives work only when it holds >= n units (stacked groups + pending children)
set -x
L=phosphorus_mk2_b200/lib
timeout 700 python tools/sweep.py --workloads spheres,terrain_bounce,terrain_nee --steps 6 $L/libphos_cuda_base.so $L/libphos_cuda.so $L/libphos_cuda_u3.so $L/libphos_cuda_u4.so $L/libphos_cuda_u5.so $L/libphos_cuda_u3i16.so $L/libphos_cuda_never.so $L/libphos_cuda.so 2>&1 | grep -v Adding | tee gpurun_out/sweep_r2p.log
