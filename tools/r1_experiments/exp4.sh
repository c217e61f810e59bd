python tools/gpu_check.py > gpurun_out/check_fast.log 2>&1; echo "check rc=$?"; grep -c "FAIL$" gpurun_out/check_fast.log; grep "FAIL$" gpurun_out/check_fast.log | head
SFC_FAST=0 python tools/gpu_check.py > gpurun_out/check_generic.log 2>&1; echo "generic check rc=$?"; grep "FAIL$" gpurun_out/check_generic.log | head
python tools/gpu_bench.py c2c4096 rfft sizes fft2 fft1m fftn blue 2>&1 | cut -c1-118
echo "--- SFC_FAST=0"; SFC_FAST=0 python tools/gpu_bench.py c2c4096 rfft 2>&1 | cut -c1-118
