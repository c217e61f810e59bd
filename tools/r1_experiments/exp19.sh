for pz in 0 2; do echo "=== SFC_PIPE=$pz"; SFC_PIPE=$pz timeout 600 python tools/gpu_bench.py fft1m blue sizes 2>&1 | grep -v "batch 1 \|f32\|x16 \|x32 \|x64 \|x128 " | cut -c1-112; done
SFC_L2_CHUNK_MB=0 SFC_PIPE=2 SFC_LIB_PATH=$PWD/build_phase/libscirs2_fft_cuda.so python tools/phase.py 2>&1 | head -12
