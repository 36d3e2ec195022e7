# round 2, call k (2 GPUs): the N-rank frame with the library's film reduce; the default bench line at N = 2
set -x
nvidia-smi -L
( time timeout 400 python -m pytest tests/test_gpu_multi.py tests/test_gpu_integration.py -m gpu -q -x --tb=short ) 2>&1 | grep -v "^[0-9]*, $\|Adding material" | tail -12 | tee gpurun_out/pytest_gpu_r2k.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 ) > gpurun_out/bench_n2_r2k.json 2> gpurun_out/bench_n2_r2k.err
tail -c 3000 gpurun_out/bench_n2_r2k.json; tail -8 gpurun_out/bench_n2_r2k.err
