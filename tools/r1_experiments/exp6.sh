python tools/gpu_check.py > gpurun_out/check6.log 2>&1; echo "check rc=$?"; grep "FAIL$" gpurun_out/check6.log | head
python tools/gpu_bench.py c2c4096 fft2 fft1m blue fftn 2>&1 | cut -c1-118
