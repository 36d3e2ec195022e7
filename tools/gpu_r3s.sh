# round 2, call 3s: does the L1 size matter to trace_kernel?  shared-memory carve-out forced to 100 % (L1 ~28 KB) / 75 % / 66 % / left to the driver;
# and a 9-entry stack (7 CTAs then fit a 132 KB carve-out: L1 124 KB instead of 92 KB) with the carve-out asked for explicitly
L=phosphorus_mk2_b200/lib
timeout 900 python tools/sweep.py --workloads spheres,terrain_bounce,terrain_nee --steps 8 $L/libphos_cuda.so $L/libphos_cuda.so:PHOS_TRACE_CARVEOUT=100 $L/libphos_cuda.so:PHOS_TRACE_CARVEOUT=75 $L/libphos_cuda.so:PHOS_TRACE_CARVEOUT=66 $L/libphos_cuda_s9.so $L/libphos_cuda_s9.so:PHOS_TRACE_CARVEOUT=57 $L/libphos_cuda_s9.so:PHOS_TRACE_CARVEOUT=50 $L/libphos_cuda.so 2>&1 | grep -v Adding | tee gpurun_out/sweep_r3s.log
