# Round 2, sixteenth GPU call (1 GPU): in-place middle stage (one barrier of three removed) — parity, racecheck, sustained A/B against
# the same sources built with -DSFC_INPLACE_MID=0 (build_ab).
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
cat > /tmp/san5.py <<'PY'
import sys; sys.path.insert(0, '.')
import numpy as np
from scirs_b200 import FftPlan
rng = np.random.default_rng(0)
def c(*s): return rng.standard_normal(s) + 1j * rng.standard_normal(s)
w = 0.0
for n in (512, 1024, 2048, 4096):
    a = c(6, n); r = FftPlan([6, n], [1]).execute(a).reshape(6, n); w = max(w, np.abs(r - np.fft.fft(a, axis=1)).max())
    a = c(2, n, 16); r = FftPlan([2, n, 16], [1]).execute(a).reshape(2, n, 16); w = max(w, np.abs(r - np.fft.fft(a, axis=1)).max())
    x = rng.standard_normal((4, 2 * n)); r = FftPlan([4, 2 * n], [1], "r2c").execute(x).reshape(4, n + 1); w = max(w, np.abs(r - np.fft.rfft(x, axis=1)).max())
a = c(2, 5000); r = FftPlan([2, 5000], [1]).execute(a).reshape(2, 5000); w = max(w, np.abs(r - np.fft.fft(a, axis=1)).max())
print("in-place middle stage workload: worst abs error", w)
PY
compute-sanitizer --tool racecheck --error-exitcode 9 python /tmp/san5.py > gpurun_out/sanitize4_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/sanitize4_racecheck.log
for i in 1 2; do
python tools/ab_headline.py 65536 4096
SFC_LIB_PATH=$PWD/build_ab/libscirs2_fft_cuda.so python tools/ab_headline.py 65536 4096
done
python tools/ab_headline.py 131072 2048
SFC_LIB_PATH=$PWD/build_ab/libscirs2_fft_cuda.so python tools/ab_headline.py 131072 2048
python tools/ab_headline.py 65536 4096 r2c
SFC_LIB_PATH=$PWD/build_ab/libscirs2_fft_cuda.so python tools/ab_headline.py 65536 4096 r2c
python tools/ab_headline.py 262144 1024
SFC_LIB_PATH=$PWD/build_ab/libscirs2_fft_cuda.so python tools/ab_headline.py 262144 1024
python tools/ab_headline.py 32 1000003
SFC_LIB_PATH=$PWD/build_ab/libscirs2_fft_cuda.so python tools/ab_headline.py 32 1000003
