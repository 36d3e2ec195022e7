# round 2, call 3g: integrate_kernel with the slot-index loads in one round trip (int1), + shadow direction / light terms in the
# batch of the hit fields (default), at 3 / 4 / 5 blocks per SM; int0 = the previous kernel
L=$PWD/phosphorus_mk2_b200/lib
run() { PHOS_CUDA_LIB=$3 python bench.py --render --workload $1 --spp 64 --depth 8 --steps 4 --warmup 2 2>/dev/null | tail -1 | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('$1 $2', round(j['ms_per_frame'],2), 'ms', round(j['value']/1e6,1), 'Msamples/s', j['image_mean'])"; }
for rep in 1 2; do
for v in int0 "" int1 int2b3 int2b5 int1b5; do
  lib=$L/libphos_cuda${v:+_$v}.so
  run cornell "${v:-default}" $lib
done
done
for v in int0 "" int1 int2b3; do
  lib=$L/libphos_cuda${v:+_$v}.so
  run terrain_ggx "${v:-default}" $lib
done
