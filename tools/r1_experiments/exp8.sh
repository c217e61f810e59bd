python tools/gpu_check.py > gpurun_out/check8.log 2>&1; echo "check rc=$?"; grep "FAIL$" gpurun_out/check8.log | head
for g in 0 1 2; do echo "--- SFC_GROUPS=$g"; SFC_GROUPS=$g python tools/gpu_bench.py fftn1024 fft1m blue 2>&1 | cut -c1-118; done
