export SFC_BLUE3_MIN=32768
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "bluestein or golden_lengths" 2>&1 | tail -3
python - <<'PY'
import numpy as np, sys
sys.path.insert(0, '.')
from scirs_b200 import FftPlan
rng = np.random.default_rng(1)
for n, b in ((20011, 5), (100003, 3), (300007, 2), (1000003, 2)):
    a = rng.standard_normal((b, n)) + 1j * rng.standard_normal((b, n))
    p = FftPlan([b, n], [1]); d = p.describe()
    got = p.execute(a).reshape(b, n); ref = np.fft.fft(a, axis=1)
    inv = FftPlan([b, n], [1], "c2c", "f64", False, 1.0 / n).execute(got).reshape(b, n)
    print(n, "Bluestein-3" in d, np.linalg.norm(got - ref) / np.linalg.norm(ref), np.linalg.norm(inv - a) / np.linalg.norm(a))
PY
for m in 0 2097152 1048576; do echo "== SFC_BLUE3_MIN=$m"; SFC_BLUE3_MIN=$m python tools/gpu_bench.py blue 2>&1 | cut -c1-112; done
SFC_BLUE3_MIN=65536 python tools/exp13.py 2>&1 | grep "rows 1342x100003" | cut -c1-100; SFC_BLUE3_MIN=0 python tools/exp13.py 2>&1 | grep "rows 1342x100003" | cut -c1-100
