# round 2, call 3c: eager commit of the next node (+ prefetch of it / of the node's first leaf triangles), L1 eviction hints on node / triangle loads
set -x
L=phosphorus_mk2_b200/lib
timeout 900 python tools/sweep.py --workloads spheres,terrain_bounce,terrain_nee --steps 8 $L/libphos_cuda_orig.so $L/libphos_cuda.so $L/libphos_cuda_eag.so $L/libphos_cuda_eagp1.so $L/libphos_cuda_eagp2.so $L/libphos_cuda_eagp3.so $L/libphos_cuda_eagp1t.so $L/libphos_cuda_nl1.so $L/libphos_cuda_tl2.so $L/libphos_cuda_nl1tl3.so $L/libphos_cuda_orig.so $L/libphos_cuda_eag.so 2>&1 | grep -v Adding | tee gpurun_out/sweep_r3c.log
