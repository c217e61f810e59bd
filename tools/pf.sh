for pf in 0 1 2 4; do echo "--- SFC_PREFETCH=$pf"; SFC_PREFETCH=$pf python tools/gpu_bench.py c2c4096 rfft 2>&1 | cut -c1-140; done
SFC_PREFETCH=1 python tools/gpu_bench.py sizes | cut -c1-140
