#!/bin/bash
# 8-GPU box: weak-scaling ray query and config 4 tile-partitioned (the short records)
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 8 "${@:2}" 2>/dev/null | grep '^{' ; }
run 29511 --no-cpu-baseline > gpurun_out/scale8_rays_r1n.json; cut -c1-330 gpurun_out/scale8_rays_r1n.json
run 29512 --render --workload terrain_ggx --spp 64 --depth 8 --steps 10 > gpurun_out/scale8_config4_r1n.json; cut -c1-300 gpurun_out/scale8_config4_r1n.json
