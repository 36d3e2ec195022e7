run() { python bench.py --render --workload $1 --spp 64 --depth 8 --steps 4 --warmup 2 2>/dev/null | tail -1 | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('$1 $2', round(j['ms_per_step'],2), 'ms', round(j['value']/1e6,1), 'Msamples/s', j['image_mean'])"; }
run cornell first; run cornell second; run cornell third; run terrain_ggx x; run cornell fourth
python -m pytest tests/test_gpu_render.py -m gpu -x -q 2>&1 | tail -2
