#!/bin/bash
# multi-GPU box: ray-query bench (value + e2e) at N = 1, 2, 4, 8
python bench.py --no-cpu-baseline 2>/dev/null | grep '^{' | python -c "import json,sys; j=json.loads(sys.stdin.read()); print(1, '%.0f'%j['value'], 'e2e %.0f'%j['e2e']['value'], j['e2e'].get('host_cpus_bound'))"
for n in 2 4 8; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --no-cpu-baseline 2>/dev/null | grep '^{' | tee gpurun_out/scale${n}_rays_r1h.json | python -c "import json,sys; j=json.loads(sys.stdin.read()); print($n, '%.0f'%j['value'], 'e2e %.0f'%j['e2e']['value'], j['e2e'].get('host_cpus_bound'))"
done
nvidia-smi topo -m 2>/dev/null | head -14
