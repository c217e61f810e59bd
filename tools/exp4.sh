python tools/gpu_check.py > gpurun_out/check_pad.log 2>&1; echo "check rc=$?"; grep -c "FAIL$" gpurun_out/check_pad.log; grep "FAIL$" gpurun_out/check_pad.log | head
python tools/gpu_bench.py c2c4096 rfft sizes fft2 fft1m fftn blue 2>&1 | cut -c1-118
