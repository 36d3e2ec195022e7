# round 2, call 3n: per-warp timeline of the final kernel's launches (-DPHOS_TAIL_PROBE), without and with lazy staging (last 8 chunks per warp)
L=phosphorus_mk2_b200/lib
rm -f /tmp/probe_*.bin
for v in probe probelz; do
  echo "== $v"
  python tools/sweep.py --workloads terrain_bounce --steps 4 $L/libphos_cuda_$v.so:PHOS_TAIL_PROBE_FILE=/tmp/probe_$v.bin 2>&1 | grep -v Adding
  python tools/tail_probe.py /tmp/probe_$v.bin 400000 | tail -3
done
