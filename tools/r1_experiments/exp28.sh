python - <<'PY'
import numpy as np, sys, os
sys.path.insert(0, '.')
os.environ["SFC_GPIPE"] = "1"; os.environ["SFC_GPIPE_MIN_TILES"] = "1"
from scirs_b200 import FftPlan
rng = np.random.default_rng(2)
for shape in ([700, 4096], [1301, 4096], [2400, 2048], [4800, 1024], [3, 4096]):
    a = rng.standard_normal(shape) + 1j * rng.standard_normal(shape)
    p = FftPlan(shape, [1]); d = p.describe()
    got = p.execute(a).reshape(shape); ref = np.fft.fft(a, axis=1)
    inv = FftPlan(shape, [1], "c2c", "f64", False, 1.0 / shape[1]).execute(got).reshape(shape)
    print(shape, "group-pipelined" in d, np.linalg.norm(got - ref) / np.linalg.norm(ref), np.linalg.norm(inv - a) / np.linalg.norm(a), flush=True)
PY
for g in 0 1; do echo "== SFC_GPIPE=$g"; SFC_GPIPE=$g timeout 300 python tools/gpu_bench.py c2c4096 sizes 2>&1 | grep -v "f32\|x16 \|x32 \|x64 \|x128 \|x256 \|x512 \|x8192" | cut -c1-112; done
