#!/bin/bash
# GPU box: warp execution efficiency + duration of the traversal kernel for several library variants
# on the config-3 bounce rays (ncu, two metrics only).  usage: tools/eff_probe.sh lib1.so lib2.so ...
for lib in "$@"; do
  ncu --metrics smsp__thread_inst_executed_per_inst_executed.ratio,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active \
      --clock-control none -k regex:trace_kernel -s 4 -c 1 --csv python tools/sweep.py --workloads terrain_bounce --steps 2 "$lib" 2>/dev/null \
    | grep -E "trace_kernel" | awk -F'","' -v L="$(basename $lib)" '{print L, $(NF-2), $NF}' | tr -d '"'
done
