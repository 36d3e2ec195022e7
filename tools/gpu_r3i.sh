# round 2, call 3i: lazy staging, window = the last 2 / 4 / 6 / 8 / 16 chunks per warp / the whole stream (lz1000)
set -x
L=phosphorus_mk2_b200/lib
timeout 900 python tools/sweep.py --workloads spheres,terrain_bounce,terrain_nee --steps 8 $L/libphos_cuda.so $L/libphos_cuda_lz2.so $L/libphos_cuda_lz4.so $L/libphos_cuda_lz6.so $L/libphos_cuda_lz8.so $L/libphos_cuda_lz16.so $L/libphos_cuda_lz1000.so $L/libphos_cuda.so $L/libphos_cuda_lz4.so $L/libphos_cuda_lz8.so 2>&1 | grep -v Adding | tee gpurun_out/sweep_r3i.log
