# compute-sanitizer (memcheck + racecheck) over the kernels added in round 2: power-of-three tiles (every length, rows,
# strided lanes and the mixed mapping of the four-step pass B, with the conflict-free lane bases), fused DCT-IV / DST-IV,
# and the tensor-map late-prefetch flavour of the narrow strided tiles.
cat > /tmp/san4.py <<'PY'
import sys; sys.path.insert(0, '.')
import numpy as np
import scirs_b200 as sb
from scirs_b200 import FftPlan
rng = np.random.default_rng(0)
def c(*s): return rng.standard_normal(s) + 1j * rng.standard_normal(s)
worst = 0.0
for shape, axes in (([5, 9], [1]), ([4, 27], [1]), ([7, 81], [1]), ([3, 243], [1]), ([2, 729], [1]), ([2, 2187], [1]), ([3, 2187], [1]),
                    ([81, 6], [0]), ([3, 243, 18], [1]), ([2, 729, 6], [1]), ([2, 27, 162], [1]), ([2, 6561], [1]), ([1, 19683], [1]),
                    ([27, 9, 4], [0, 1]), ([1, 177147], [1])):
    x = c(*shape)
    for fwd in (True, False):
        y = FftPlan(shape, axes, "c2c", "f64", fwd).execute(x).reshape(shape)
        ref = np.fft.fftn(x, axes=axes) if fwd else np.fft.ifftn(x, axes=axes) * np.prod([shape[a] for a in axes])
        worst = max(worst, np.linalg.norm(y - ref) / np.linalg.norm(ref))
x = c(2, 729).astype(np.complex64)
y = FftPlan([2, 729], [1], "c2c", "f32", True).execute(x).reshape(2, 729)
print("r3 worst rel-L2 f64", worst, "f32", np.linalg.norm(y - np.fft.fft(x.astype(np.complex128), axis=1)) / np.linalg.norm(np.fft.fft(x, axis=1)))
for t in (4,):
    v = rng.standard_normal((3, 256)); sb.dct(v, t); sb.dst(v, t); sb.dctn(rng.standard_normal((4, 128, 6)), t, None, [1])
a = c(2, 1 << 16)
r = FftPlan([2, 1 << 16], [1]).execute(a).reshape(2, 1 << 16)
print("four-step 2^16 (narrow strided tiles, tensor-map prefetch when enabled):", np.abs(r - np.fft.fft(a, axis=1)).max())
print("sanitizer workload done")
PY
for tool in memcheck racecheck; do
  SFC_PIPE_LATE_MIN_TILES=1 compute-sanitizer --tool $tool --error-exitcode 9 python /tmp/san4.py > gpurun_out/sanitize3_$tool.log 2>&1; echo "$tool rc=$?"; tail -4 gpurun_out/sanitize3_$tool.log
done
