# round 2, call w: in-frame shadow rays without the light's ids (A/B through PHOS_SHADE_EXPORT_IDS)
set -x
( timeout 400 python -m pytest tests/test_gpu_render.py tests/test_gpu_trace.py tests/test_gpu_integration.py -m gpu -q -x --tb=short ) 2>&1 | tail -3 | tee gpurun_out/pytest_gpu_r2w.log
for rep in 1 2; do for m in "" 1; do for w in cornell terrain_ggx; do
  PHOS_SHADE_EXPORT_IDS=$m; [ -z "$m" ] && unset PHOS_SHADE_EXPORT_IDS || export PHOS_SHADE_EXPORT_IDS
  timeout 300 python bench.py --render --workload $w --spp 64 --depth 8 --steps 6 --warmup 2 2>/dev/null | tail -1 | cut -c1-120 | sed "s/^/export_ids=$m $w /" | tee -a gpurun_out/render_r2w.log
done; done; done
