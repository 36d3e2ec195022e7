# round 2, call 3k: the default bench line at 8 GPUs with the final kernels (eager commit, batched integrate loads, 128 Ki pipeline chunks)
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 3 2> gpurun_out/bench_n8_r3k.err | grep '^{' > gpurun_out/bench_n8_r3k.json
cut -c1-400 gpurun_out/bench_n8_r3k.json; tail -3 gpurun_out/bench_n8_r3k.err
