# round 2, call 3a: traversal CTAs per SM inside a frame (room for the other wavefront's shading kernels)
set -x
for m in 0 6 5 4; do for w in terrain_ggx cornell; do
  PHOS_RENDER_TRACE_BLOCKS=$m timeout 300 python bench.py --render --workload $w --spp 64 --depth 8 --steps 6 --warmup 2 2>/dev/null | tail -1 | cut -c1-120 | sed "s/^/trace_blocks=$m $w /" | tee -a gpurun_out/render_r3a.log
done; done
