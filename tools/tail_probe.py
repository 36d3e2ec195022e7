#!/usr/bin/env python3
"""Read the per-warp timeline a -DPHOS_TAIL_PROBE build of the library appends to $PHOS_TAIL_PROBE_FILE
(one record per trace_kernel launch: n_warps, n_rays, then 5 words per warp: start, stream-dry, end in
globaltimer ns, loop iterations after the stream ran dry, lane-iterations with work after it) and print where
a launch spends its time: ramp-up, the part with rays left to claim, and the drain after the stream ran dry.
Usage: python tools/tail_probe.py FILE [min_rays]"""
import sys

import numpy as np


def main():
    raw = np.fromfile(sys.argv[1], dtype=np.uint64)
    min_rays = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
    o = 0
    k = 0
    while o < len(raw):
        nw, n = int(raw[o]), int(raw[o + 1])
        rec = raw[o + 2:o + 2 + 5 * nw].reshape(nw, 5).astype(np.int64)
        o += 2 + 5 * nw
        k += 1
        rec = rec[rec[:, 2] > 0]
        if n < min_rays or len(rec) == 0:
            continue
        t0 = rec[:, 0].min()
        start, dry, end = rec[:, 0] - t0, rec[:, 1] - t0, rec[:, 2] - t0
        has_dry = rec[:, 1] > 0
        first_dry = dry[has_dry].min() if has_dry.any() else -1
        total = end.max()
        q = np.percentile(end, [1, 10, 50, 90, 99, 100]) / 1e3
        it, ln = rec[:, 3].sum(), rec[:, 4].sum()
        print(f"launch {k:3d}: {n:9d} rays {len(rec):5d} warps  total {total/1e3:7.1f} us  last warp start {start.max()/1e3:6.1f} us  "
              f"stream dry at {first_dry/1e3:7.1f} us ({100.0*first_dry/total:4.1f} %)  drain {(total-first_dry)/1e3:6.1f} us  "
              f"warp end pct[1,10,50,90,99,100] = {np.array2string(q, precision=1, floatmode='fixed')}  "
              f"lanes/iteration after dry {ln/max(it,1):4.1f}  iterations after dry per warp {it/len(rec):6.1f}")


if __name__ == "__main__":
    main()
