#!/bin/bash
# 2-GPU sanity record after kernel / wavefront changes: weak-scaling ray query and config 4 tile-partitioned
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 "${@:2}" 2>/dev/null | grep '^{' ; }
run 29511 --no-cpu-baseline > gpurun_out/scale2_rays_r1n.json; cut -c1-330 gpurun_out/scale2_rays_r1n.json
run 29512 --render --workload terrain_ggx --spp 64 --depth 8 --steps 6 > gpurun_out/scale2_config4_r1n.json; cut -c1-300 gpurun_out/scale2_config4_r1n.json
