# Round 2, seventeenth GPU call (1 GPU): in-place middle stage also in the fused real-transform flavours — parity, racecheck, A/B.
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
cat > /tmp/san5.py <<PY
import sys; sys.path.insert(0, ".")
import numpy as np
from scirs_b200 import FftPlan
rng = np.random.default_rng(0)
w = 0.0
for n in (1024, 2048, 4096, 8192):
    x = rng.standard_normal((4, n)); r = FftPlan([4, n], [1], "r2c").execute(x).reshape(4, n // 2 + 1); w = max(w, np.abs(r - np.fft.rfft(x, axis=1)).max())
    s = np.fft.rfft(x, axis=1); r = FftPlan([4, n], [1], "c2r", "f64", False, 1.0 / n).execute(s).reshape(4, n); w = max(w, np.abs(r - x).max())
print("real flavours, worst abs error", w)
PY
compute-sanitizer --tool racecheck --error-exitcode 9 python /tmp/san5.py > gpurun_out/sanitize4_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/sanitize4_racecheck.log
for k in r2c c2r; do
python tools/ab_headline.py 65536 4096 $k
SFC_LIB_PATH=$PWD/build_ab/libscirs2_fft_cuda.so python tools/ab_headline.py 65536 4096 $k
python tools/ab_headline.py 65536 4096 $k f32
SFC_LIB_PATH=$PWD/build_ab/libscirs2_fft_cuda.so python tools/ab_headline.py 65536 4096 $k f32
done
python tools/ab_headline.py 65536 4096 c2c f32
SFC_LIB_PATH=$PWD/build_ab/libscirs2_fft_cuda.so python tools/ab_headline.py 65536 4096 c2c f32
