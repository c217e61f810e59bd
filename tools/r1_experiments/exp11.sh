python tools/gpu_check.py > gpurun_out/check11.log 2>&1; echo "check rc=$?"; grep "FAIL$" gpurun_out/check11.log | head -5; grep "f32" gpurun_out/check11.log | sort -t= -k2 -g | tail -2
python tools/gpu_bench.py c2c4096 rfft | grep f32 | cut -c1-118
