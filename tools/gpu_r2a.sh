# round 2, call a: GPU suite as it stands + per-warp timeline of a launch (tail probe)
set -x
( time python -m pytest tests -m gpu -x -q ) 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_r2a.log
rm -f /tmp/probe_*.bin
L=phosphorus_mk2_b200/lib/libphos_cuda_probe.so
python tools/sweep.py --workloads spheres --steps 4 $L:PHOS_TAIL_PROBE_FILE=/tmp/probe_spheres.bin 2>&1 | grep -v Adding | tee gpurun_out/probe_r2a.log
python tools/sweep.py --workloads terrain_bounce --steps 4 $L:PHOS_TAIL_PROBE_FILE=/tmp/probe_bounce.bin 2>&1 | grep -v Adding | tee -a gpurun_out/probe_r2a.log
python tools/sweep.py --workloads terrain_nee --steps 4 $L:PHOS_TAIL_PROBE_FILE=/tmp/probe_nee.bin 2>&1 | grep -v Adding | tee -a gpurun_out/probe_r2a.log
for w in spheres bounce nee; do echo "== $w"; python tools/tail_probe.py /tmp/probe_$w.bin 400000 | tail -4; done 2>&1 | tee -a gpurun_out/probe_r2a.log
