python -m pytest tests/test_gpu_render.py tests/test_gpu_integration.py tests/test_codec.py -m gpu -x -q 2>&1 | tail -3
run() { python bench.py --render --workload $1 --spp 64 --depth $2 --steps 3 --warmup 2 2>/dev/null | tail -1 | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('$1 depth $2 paths ${PHOS_WAVEFRONT_PATHS:-default}', round(j['ms_per_step'],2), 'ms', round(j['value']/1e6,1), 'Msamples/s', j['gpu_launches'], j['image_mean'])"; }
for d in 1 2 4 8; do run terrain_ggx $d; done
for p in 8388608 16777216 33554432 67108864; do PHOS_WAVEFRONT_PATHS=$p run terrain_ggx 8; done
run cornell 8
for p in 8388608 16777216; do PHOS_WAVEFRONT_PATHS=$p run cornell 8; done
