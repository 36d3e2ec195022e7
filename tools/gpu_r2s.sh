# round 2, call s: light terms of the next-event estimate computed in shade_nee (A/B against the previous shading kernels)
set -x
L=phosphorus_mk2_b200/lib
( timeout 400 python -m pytest tests/test_gpu_render.py tests/test_gpu_integration.py -m gpu -q -x --tb=short ) 2>&1 | tail -3 | tee gpurun_out/pytest_gpu_r2s.log
for rep in 1 2; do for v in "" _prevshade; do for w in cornell terrain_ggx; do
  PHOS_CUDA_LIB=$PWD/$L/libphos_cuda$v.so timeout 300 python bench.py --render --workload $w --spp 64 --depth 8 --steps 6 --warmup 2 2>/dev/null | tail -1 | cut -c1-120 | sed "s/^/lib=$v $w /" | tee -a gpurun_out/render_r2s.log
done; done; done
timeout 700 python tools/sweep.py --workloads spheres,terrain_bounce,terrain_nee --steps 8 $L/libphos_cuda_base.so $L/libphos_cuda.so $L/libphos_cuda_base.so $L/libphos_cuda.so 2>&1 | grep -v Adding | tee gpurun_out/sweep_r2s.log
