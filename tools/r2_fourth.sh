# Round 2, fourth GPU call (1 GPU): copy power ceiling; tensor-map column flavour parity + timing; DCT-IV default on.
python tools/copy_power.py
echo "=== tensor-map column tiles (SFC_PIPE_LATE=3)"
SFC_PIPE_LATE=3 SFC_PIPE_LATE_MIN_TILES=1 timeout 600 python -m pytest tests/test_gpu_knobs.py -x -q -m gpu -k "defaults" 2>&1 | tail -5
SFC_PIPE_LATE=3 SFC_PIPE_LATE_MIN_TILES=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_random.py -x -q -m gpu 2>&1 | tail -5
echo "--- timings default"
python tools/gpu_bench.py fftn fft1m blue 2>&1 | tail -9
echo "--- timings SFC_PIPE_LATE=3"
SFC_PIPE_LATE=3 python tools/gpu_bench.py fftn fft1m blue 2>&1 | tail -9
SFC_PIPE_LATE=3 python - <<'PY'
import sys; sys.path.insert(0, '.')
from scirs_b200 import FftPlan
for sh, ax in (([512,512,512],[0,1,2]), ([64, 1<<20],[1]), ([32, 1000003],[1])):
    print(FftPlan(sh, ax).describe())
PY
echo "=== consumers with the fused DCT-IV default"
timeout 900 python -m pytest tests/test_gpu_consumers.py tests/test_gpu_experimental.py -x -q -m gpu 2>&1 | tail -5
