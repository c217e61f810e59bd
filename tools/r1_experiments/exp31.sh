# A/B: stage twiddles from the table (default build) against the product tree (build_x, -DSFC_TW_LOAD=0)
timeout 100 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "golden or pow2 or rfft or four_step" 2>&1 | tail -2
for r in 1 2; do
echo "== table"; timeout 60 python tools/gpu_bench.py c2c4096 rfft 2>&1 | cut -c1-112
echo "== tree";  SFC_LIB_PATH=build_x/libscirs2_fft_cuda.so timeout 60 python tools/gpu_bench.py c2c4096 rfft 2>&1 | cut -c1-112
done
