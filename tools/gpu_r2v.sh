# round 2, call v: write-back skips groups of unhit closest-hit rays (host-pointer query, e2e)
set -x
( timeout 300 python -m pytest tests/test_gpu_trace.py tests/test_gpu_render.py -m gpu -q -x --tb=short ) 2>&1 | tail -3 | tee gpurun_out/pytest_gpu_r2v.log
for rep in 1 2 3; do for m in 0 1; do
  PHOS_E2E_SKIP_UNCHANGED=$m timeout 300 python bench.py --only headline --no-cpu-baseline --steps 10 --warmup 3 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('skip=$m', d['value'], d['e2e'])" | tee -a gpurun_out/e2e_r2v.log
done; done
