# round 2, call 3d: eager commit + has_ray-free predicates (eag2), lazy staging near the end of the stream, vote / refill / sharing knobs re-swept on the eager loop
set -x
L=phosphorus_mk2_b200/lib
timeout 1200 python tools/sweep.py --workloads spheres,terrain_bounce,terrain_nee --steps 8 $L/libphos_cuda_orig.so $L/libphos_cuda_eag.so $L/libphos_cuda_eag2.so $L/libphos_cuda_eagl1.so $L/libphos_cuda_eagl2.so $L/libphos_cuda_eagl3.so $L/libphos_cuda_eagl2d.so $L/libphos_cuda_eagl4q.so $L/libphos_cuda_eagb2.so $L/libphos_cuda_eagb4.so $L/libphos_cuda_eagr4.so $L/libphos_cuda_eagr8.so $L/libphos_cuda_eags8.so $L/libphos_cuda_eags16.so $L/libphos_cuda_eag.so $L/libphos_cuda_eag2.so 2>&1 | grep -v Adding | tee gpurun_out/sweep_r3d.log
