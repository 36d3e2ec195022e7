# round 2, call 3r: integrate_kernel at 2 blocks per SM (no register cap worth the name) against the shipped 3
L=$PWD/phosphorus_mk2_b200/lib
run() { PHOS_CUDA_LIB=$3 python bench.py --render --workload $1 --spp 64 --depth 8 --steps 4 --warmup 2 2>/dev/null | tail -1 | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('$1 $2', round(j['ms_per_frame'],2), 'ms', round(j['value']/1e6,1), 'Msamples/s', j['image_mean'])"; }
for rep in 1 2; do for v in "" int2b2; do run cornell "${v:-default}" $L/libphos_cuda${v:+_$v}.so; done; done
for v in "" int2b2; do run terrain_ggx "${v:-default}" $L/libphos_cuda${v:+_$v}.so; done
