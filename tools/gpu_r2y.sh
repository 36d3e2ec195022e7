# round 2, call y: 1..4 wavefronts on as many streams
set -x
( timeout 400 python -m pytest tests/test_gpu_render.py tests/test_gpu_integration.py tests/test_gpu_multi.py -m gpu -q -x --tb=short ) 2>&1 | tail -3 | tee gpurun_out/pytest_gpu_r2y.log
for rep in 1 2; do for m in 1 2 3 4; do for w in terrain_ggx cornell; do
  PHOS_WAVEFRONTS=$m timeout 300 python bench.py --render --workload $w --spp 64 --depth 8 --steps 6 --warmup 2 2>/dev/null | tail -1 | cut -c1-120 | sed "s/^/wavefronts=$m $w /" | tee -a gpurun_out/render_r2y.log
done; done; done
