# round 2, call 3p: the default bench line at 2 GPUs with the final kernels
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 3 2> gpurun_out/bench_n2_r3p.err | grep '^{' > gpurun_out/bench_n2_r3p.json
cut -c1-300 gpurun_out/bench_n2_r3p.json; tail -2 gpurun_out/bench_n2_r3p.err
