# compute-sanitizer (memcheck + racecheck) over a compact set of plans covering every kernel flavour
cat > /tmp/san.py <<'PY'
import sys; sys.path.insert(0, '.')
import numpy as np
import scirs_b200 as sb
from scirs_b200 import FftPlan
rng = np.random.default_rng(0)
def c(*s): return rng.standard_normal(s) + 1j * rng.standard_normal(s)
FftPlan([3, 4096], [1]).execute(c(3, 4096))          # fast c2c rows (partial tile -> generic)
FftPlan([4, 4096], [1]).execute(c(4, 4096))          # fast c2c rows
FftPlan([2, 512, 24], [1]).execute(c(2, 512, 24))    # column tiles
FftPlan([2, 1 << 14], [1]).execute(c(2, 1 << 14))    # four-step
FftPlan([2, 1000], [1]).execute(c(2, 1000))          # Bluestein single kernel
FftPlan([2, 5000], [1]).execute(c(2, 5000))          # Bluestein three-pass
FftPlan([4, 4096], [1], "r2c").execute(rng.standard_normal((4, 4096)))   # mirrored r2c
FftPlan([4, 4096], [1], "c2r").execute(c(4, 2049))                       # mirrored c2r
FftPlan([16, 1024], [1], "r2c").execute(rng.standard_normal((16, 1024))) # staged r2c
FftPlan([16, 1024], [1], "c2r").execute(c(16, 513))                      # staged c2r
FftPlan([6, 10, 14], None, "c2r", scale=1.0).execute(c(6, 10, 8))        # Hermitian fill path
sb.fftn(rng.standard_normal((5, 6, 7)), [8, 8, 8], [2, 0], "ortho")      # pad/convert copy
FftPlan([8, 64, 4], [1], scatter_parts=0).execute(c(8, 64, 4))
print("sanitizer workload done")
PY
for tool in memcheck racecheck; do
  compute-sanitizer --tool $tool --error-exitcode 9 python /tmp/san.py > gpurun_out/sanitize_$tool.log 2>&1; echo "$tool rc=$?"; tail -3 gpurun_out/sanitize_$tool.log
done
