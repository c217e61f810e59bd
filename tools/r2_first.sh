# Round 2, first GPU call: full suite, experimental paths, timings, default bench.
python -m pytest tests -x -q -m gpu 2>&1 | tail -5
SFC_TEST_EXPERIMENTAL=1 python -m pytest tests/test_gpu_experimental.py -x -q -m gpu 2>&1 | tail -25
SFC_FFT2_TILE2D=0 python tools/exp32.py 2>&1 | tail -40
SFC_FFT2_TILE2D=1 python tools/exp32.py 2>&1 | tail -24
python tools/gpu_bench.py all 2>&1 | tail -60
python bench.py 2>&1 | tail -3
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv
