# round 2, call m (8 GPUs): the default bench line at N = 8 (config 4 tile-partitioned + config 5 sample-partitioned with the
# library's film reduce), two wavefronts at 8 GPUs, and the host-pointer query with copy-engine write-back instead of
# zero-copy stores (is the N-GPU e2e ceiling host DRAM or the store pattern?)
set -x
nvidia-smi -L | wc -l
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port"
( time timeout 600 $T 29521 bench.py --gpus 8 --steps 10 --warmup 3 ) > gpurun_out/bench_n8_r2m.json 2> gpurun_out/bench_n8_r2m.err; tail -c 1200 gpurun_out/bench_n8_r2m.json; tail -3 gpurun_out/bench_n8_r2m.err
( PHOS_WAVEFRONTS=2 timeout 400 $T 29522 bench.py --gpus 8 --only config4 ) 2>/dev/null | tail -1 | cut -c1-900 | tee gpurun_out/bench_n8_wf2_r2m.json
( PHOS_E2E_SPARSE=0 timeout 400 $T 29523 bench.py --gpus 8 --only headline --steps 10 --warmup 3 ) 2>/dev/null | tail -1 | cut -c1-1500 | tee gpurun_out/bench_n8_dense_r2m.json
( time timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q -x --tb=short ) 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_r2m.log
