#!/bin/bash
# GPU box: path-traced samples/s of bench.py --render for several library variants ($PHOS_CUDA_LIB)
for lib in "$@"; do for wl in cornell terrain_ggx; do
  PHOS_CUDA_LIB=$PWD/$lib python bench.py --render --workload $wl --spp 64 --depth 8 --steps 5 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print('$(basename $lib) $wl %.1f Msamples/s  %.2f ms/step' % (j['value']/1e6, j['ms_per_step']))"
done; done
