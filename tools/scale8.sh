#!/bin/bash
# 8-GPU box: the three multi-GPU records (weak-scaling ray query, config 4 tile-partitioned, config 5 sample-partitioned)
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 8 "${@:2}" 2>/dev/null | grep '^{' ; }
run 29511 --no-cpu-baseline > gpurun_out/scale8_rays_r1g.json; cut -c1-400 gpurun_out/scale8_rays_r1g.json
run 29512 --render --workload terrain_ggx --spp 64 --depth 8 --steps 10 > gpurun_out/scale8_config4_r1g.json; cut -c1-300 gpurun_out/scale8_config4_r1g.json
run 29513 --render --workload instanced30m --spp 1024 --depth 8 --partition samples --steps 2 --warmup 1 > gpurun_out/scale8_config5_r1g.json; cut -c1-300 gpurun_out/scale8_config5_r1g.json
