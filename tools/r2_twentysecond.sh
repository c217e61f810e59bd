# Round 2, twenty-second GPU call (1 GPU): in-place middle stage in the late-prefetch (tensor-map) flavour — A/B against the same
# sources with -DSFC_INPLACE_MID_HOOK=0 (build_ab); f32 lane-group tiles; parity + racecheck of both.
python tools/ab_headline.py 64 1048576
SFC_LIB_PATH=$PWD/build_ab/libscirs2_fft_cuda.so python tools/ab_headline.py 64 1048576
python tools/ab_headline.py 32 2097152
SFC_LIB_PATH=$PWD/build_ab/libscirs2_fft_cuda.so python tools/ab_headline.py 32 2097152
python tools/ab_headline.py 524288 512 c2c f32
SFC_ROW_LANE_GROUPS=0 python tools/ab_headline.py 524288 512 c2c f32
python tools/ab_headline.py 262144 1024 c2c f32
SFC_ROW_LANE_GROUPS=0 python tools/ab_headline.py 262144 1024 c2c f32
cat > /tmp/san6.py <<'PY'
import sys; sys.path.insert(0, '.')
import numpy as np
from scirs_b200 import FftPlan
rng = np.random.default_rng(0)
def c(*s): return rng.standard_normal(s) + 1j * rng.standard_normal(s)
w = 0.0
a = c(2, 1 << 20); r = FftPlan([2, 1 << 20], [1]).execute(a).reshape(2, 1 << 20); w = max(w, np.abs(r - np.fft.fft(a, axis=1)).max())
for n in (512, 1024):
    for prec, cd in (("f64", np.complex128), ("f32", np.complex64)):
        a = c(12, n).astype(cd); r = FftPlan([12, n], [1], "c2c", prec).execute(a).reshape(12, n)
        e = np.abs(r - np.fft.fft(a.astype(np.complex128), axis=1)).max(); w = max(w, e if prec == "f64" else e * 1e-6)
print("late-prefetch in-place + lane groups workload: worst abs error (f32 scaled by 1e-6)", w)
PY
SFC_PIPE_LATE_MIN_TILES=1 compute-sanitizer --tool racecheck --error-exitcode 9 python /tmp/san6.py > gpurun_out/sanitize5_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/sanitize5_racecheck.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_knobs.py tests/test_gpu_random.py -x -q -m gpu 2>&1 | tail -3
