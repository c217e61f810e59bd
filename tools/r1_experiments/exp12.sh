python tools/gpu_check.py > gpurun_out/check12.log 2>&1; echo "check rc=$?"; grep "FAIL$" gpurun_out/check12.log | head -5; grep "n=8192" gpurun_out/check12.log | head
python tools/gpu_bench.py sizes fft2 | tail -5 | cut -c1-118
