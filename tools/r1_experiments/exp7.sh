python tools/gpu_check.py > gpurun_out/check7.log 2>&1; echo "check rc=$?"; grep "FAIL$" gpurun_out/check7.log | head
python tools/gpu_bench.py rfft 2>&1 | cut -c1-118
