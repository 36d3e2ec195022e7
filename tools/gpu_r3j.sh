# round 2, call 3j: host-pointer pipeline: chunk sizes around 128 Ki rays at 3 / 4 / 6 / 8 chunks in flight
L=$PWD/phosphorus_mk2_b200/lib
for v in "" p3 p6 p8; do
  echo "## chunks in flight: ${v:-p4 (default)}"
  PHOS_CUDA_LIB=$L/libphos_cuda${v:+_$v}.so timeout 200 python tools/e2e_chunks.py 65536,114688,122880,131072,139264,147456,262144 2>&1 | grep -v Adding
done
