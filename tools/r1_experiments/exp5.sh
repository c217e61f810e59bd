for tl in 0 4; do echo "--- SFC_COL_TL=$tl"; SFC_COL_TL=$tl python tools/gpu_bench.py fft1m blue fftn1024 2>&1 | cut -c1-118; done
