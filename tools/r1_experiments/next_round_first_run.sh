# First GPU call of the next round: the full GPU suite, then the paths that were written after this
# round's GPU budget was spent — parity first, timing second.  ~4 GPU-minutes.
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/next_round_first_run.sh > gpurun_out/first_run.log 2>&1'
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
SFC_TEST_EXPERIMENTAL=1 python -m pytest tests/test_gpu_experimental.py -x -q -m gpu 2>&1 | tail -15
SFC_FFT2_TILE2D=0 python tools/exp32.py 2>&1 | tail -40
SFC_FFT2_TILE2D=1 python tools/exp32.py 2>&1 | tail -24
