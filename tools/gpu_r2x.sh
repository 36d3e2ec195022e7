# round 2, call x: two wavefronts on two streams (PHOS_WAVEFRONTS=2) with the round-2 kernels, one GPU
set -x
for rep in 1 2; do for m in 1 2; do for w in terrain_ggx cornell; do
  PHOS_WAVEFRONTS=$m timeout 300 python bench.py --render --workload $w --spp 64 --depth 8 --steps 6 --warmup 2 2>/dev/null | tail -1 | cut -c1-120 | sed "s/^/wavefronts=$m $w /" | tee -a gpurun_out/render_r2x.log
done; done; done
for m in 1 2; do
  PHOS_WAVEFRONTS=$m timeout 400 python bench.py --render --workload instanced30m --spp 16 --depth 8 --steps 3 --warmup 1 2>/dev/null | tail -1 | cut -c1-120 | sed "s/^/wavefronts=$m instanced30m /" | tee -a gpurun_out/render_r2x.log
done
