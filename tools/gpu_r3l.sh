# round 2, call 3l: the default bench line at 4 GPUs with the final kernels
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 4 --steps 20 --warmup 3 2> gpurun_out/bench_n4_r3l.err | grep '^{' > gpurun_out/bench_n4_r3l.json
cut -c1-300 gpurun_out/bench_n4_r3l.json; tail -2 gpurun_out/bench_n4_r3l.err
