# L2-blocked multi-stream rounds: sweep chunk size / rounds in flight
run() { echo "--- $*"; env "$@" python tools/gpu_bench.py fft1m blue 2>&1 | grep -v "batch 1 " | cut -c1-112; }
run SFC_L2_CHUNK_MB=0
run SFC_L2_CHUNK_MB=17 SFC_L2_WAYS=2 SFC_L2_MAXB_MB=40
run SFC_L2_CHUNK_MB=17 SFC_L2_WAYS=3 SFC_L2_MAXB_MB=40 SFC_L2_TOTAL_MB=110
run SFC_L2_CHUNK_MB=17 SFC_L2_WAYS=4 SFC_L2_MAXB_MB=40 SFC_L2_TOTAL_MB=140
run SFC_L2_CHUNK_MB=36 SFC_L2_WAYS=2
run SFC_L2_CHUNK_MB=36 SFC_L2_WAYS=3 SFC_L2_TOTAL_MB=110
run SFC_L2_CHUNK_MB=36 SFC_L2_WAYS=4 SFC_L2_TOTAL_MB=150
run SFC_L2_CHUNK_MB=36 SFC_L2_WAYS=2 SFC_L2_MAXB_MB=80 SFC_L2_TOTAL_MB=140
run SFC_L2_CHUNK_MB=68 SFC_L2_WAYS=2 SFC_L2_TOTAL_MB=140
run SFC_L2_CHUNK_MB=68 SFC_L2_WAYS=3 SFC_L2_TOTAL_MB=210
python tools/gpu_check.py > gpurun_out/check15.log 2>&1; echo "check rc=$?"; grep -c "ok$" gpurun_out/check15.log; grep "FAIL" gpurun_out/check15.log | head -5
