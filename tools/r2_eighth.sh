# Round 2, eighth GPU call (1 GPU): re-measure the power-of-three tiles; 8192-point rows with the radix-2 row split.
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "power_of_three or 3pow13" 2>&1 | tail -3
python tools/ab_headline.py 32 1594323
SFC_R3_TL_2187=3 python tools/ab_headline.py 32 1594323
python tools/ab_headline.py 256 1594323
python tools/ab_headline.py 65536 2187
python tools/ab_headline.py 196608 729
python tools/ab_headline.py 1048576 243
python tools/ab_headline.py 2097152 81
echo "=== 8192 rows"
python tools/ab_headline.py 32768 8192
SFC_ROW_SPLIT=8192 python tools/ab_headline.py 32768 8192
SFC_ROW_SPLIT=8192 python tools/gpu_bench.py fft2 2>&1 | tail -2
SFC_ROW_SPLIT=8192 timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fft2 or golden" 2>&1 | tail -2
