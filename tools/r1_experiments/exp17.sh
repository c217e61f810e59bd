export SFC_TILE_GROUP_LOG2=3
for v in "" build_d1a0 build_d2a0 build_d3a0 build_d1a1 build_d3a3; do
  echo "=== variant ${v:-default}"
  if [ -n "$v" ]; then export SFC_LIB_PATH=$PWD/$v/libscirs2_fft_cuda.so; fi
  python tools/gpu_bench.py c2c4096 fft1m blue fftn 2>&1 | grep -v "batch 1 \|f32\|axis" | cut -c1-112
done
