#!/bin/bash
# GPU box: e2e Mrays/s of bench.py for the host-pointer pipeline variants (sparse write-back on / off, chunk sizes)
python -m pytest tests/test_gpu_trace.py -x -q -m gpu 2>&1 | tail -3
for sp in 0 1; do for ch in 65536 131072 262144 524288; do
  PHOS_E2E_SPARSE=$sp PHOS_PIPE_CHUNK=$ch python bench.py --no-cpu-baseline --steps 10 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print('sparse=$sp chunk=$ch value %.0f e2e %.0f ok=%s' % (j['value'], j['e2e']['value'], j['e2e']['matches_device_path']))"
done; done
