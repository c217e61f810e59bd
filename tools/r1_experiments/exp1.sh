for tl in 0 4; do for mb in 16 64 512; do echo "--- SFC_COL_TL=$tl SFC_WORK_MB=$mb"; SFC_COL_TL=$tl SFC_WORK_MB=$mb python tools/gpu_bench.py fft1m blue 2>&1 | cut -c1-150; done; done
for tl in 0 4; do echo "--- SFC_COL_TL=$tl"; SFC_COL_TL=$tl python tools/gpu_bench.py fftn1024 fftn 2>&1 | cut -c1-150; done
