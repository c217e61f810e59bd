# Round 2, third GPU call (1 GPU): parity of the late-prefetch flavour, then sustained A/B of the headline kernel's variants.
timeout 600 python -m pytest tests/test_gpu_knobs.py -x -q -m gpu -k "late or LATE" 2>&1 | tail -15
echo "=== sustained A/B, c2c 65536x4096 f64"
python tools/ab_headline.py
SFC_PIPE_LATE=1 python tools/ab_headline.py
SFC_PIPE=2 python tools/ab_headline.py
SFC_LIB_PATH=$PWD/build_ab/tw/libscirs2_fft_cuda.so python tools/ab_headline.py
python tools/ab_headline.py
SFC_PIPE_LATE=1 python tools/ab_headline.py
echo "=== 8192 rows"
python tools/ab_headline.py 32768 8192
SFC_PIPE_LATE=2 python tools/ab_headline.py 32768 8192
SFC_PIPE_BIG=0 python tools/ab_headline.py 32768 8192
echo "=== 2048 rows"
python tools/ab_headline.py 131072 2048
SFC_PIPE_LATE=1 python tools/ab_headline.py 131072 2048
echo "=== f32 8192 rows"
python tools/ab_headline.py 65536 8192 c2c f32
SFC_PIPE_LATE=2 python tools/ab_headline.py 65536 8192 c2c f32
echo "=== rfft (sustained)"
python tools/ab_headline.py 65536 4096 r2c f64
python tools/ab_headline.py 65536 4096 c2r f64
echo "=== fft2"
SFC_PIPE_LATE=2 python tools/gpu_bench.py fft2 2>&1 | tail -2
