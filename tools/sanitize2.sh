# compute-sanitizer (memcheck + racecheck) over the kernels added in the second half of round 1:
# generated-chirp Bluestein passes, fused-table (DCT/DST) plans, element-wise / framing / transpose kernels,
# and the (default-off) TMA-pipelined flavour.
cat > /tmp/san2.py <<'PY'
import sys, os; sys.path.insert(0, '.')
import numpy as np
import scirs_b200 as sb
from scirs_b200 import FftPlan
rng = np.random.default_rng(0)
def c(*s): return rng.standard_normal(s) + 1j * rng.standard_normal(s)
FftPlan([2, 5000], [1]).execute(c(2, 5000))                 # Bluestein three-pass with generated chirps
FftPlan([2, 20011], [1]).execute(c(2, 20011))
x = rng.standard_normal((3, 64, 5))
for t in (1, 2, 3, 4):
    sb.dctn(x, t, "ortho"); sb.idstn(x, t, None, [1])       # fused tables (pow2) and explicit pre/post passes (others)
sb.dct(rng.standard_normal(1000), 2); sb.dst(rng.standard_normal(16384), 4)
sb.dht(rng.standard_normal(100)); sb.dht2(rng.standard_normal((12, 20)), (1, 0))
sb.hfft(c(37), 50); sb.ihfft(rng.standard_normal(33), 40); sb.hilbert(rng.standard_normal(100))
s = rng.standard_normal(3000)
sb.stft(s, "hann", 100, 25, 128, boundary="reflect"); sb.spectrogram(s, nperseg=64, mode="magnitude")
sb.stft(s, "hann", 128, return_onesided=False)
sb.fft_streaming(s, None, sb.FftMode.Inverse, 1000); sb.fftn_optimized(rng.standard_normal((6, 10, 7)))
sb.fft2_efficient(rng.standard_normal((10, 12)), (16, 16), sb.FftMode.Forward, True)
print("sanitizer workload done")
PY
for tool in memcheck racecheck; do
  compute-sanitizer --tool $tool --error-exitcode 9 python /tmp/san2.py > gpurun_out/sanitize2_$tool.log 2>&1; echo "$tool rc=$?"; tail -3 gpurun_out/sanitize2_$tool.log
done
cat > /tmp/san3.py <<'PY'
import sys; sys.path.insert(0, '.')
import numpy as np
from scirs_b200 import FftPlan
rng = np.random.default_rng(0)
def c(*s): return rng.standard_normal(s) + 1j * rng.standard_normal(s)
a = c(8, 4096); r = FftPlan([8, 4096], [1]).execute(a); print(np.abs(r.reshape(8, 4096) - np.fft.fft(a, axis=1)).max())
a = c(2, 512, 16); r = FftPlan([2, 512, 16], [1]).execute(a); print(np.abs(r.reshape(2, 512, 16) - np.fft.fft(a, axis=1)).max())
print("pipelined flavour done")
PY
SFC_PIPE=1 SFC_PIPE_MIN_TILES=1 compute-sanitizer --tool memcheck --error-exitcode 9 python /tmp/san3.py > gpurun_out/sanitize2_pipe_memcheck.log 2>&1; echo "pipe memcheck rc=$?"; tail -4 gpurun_out/sanitize2_pipe_memcheck.log
SFC_PIPE=1 SFC_PIPE_MIN_TILES=1 compute-sanitizer --tool racecheck --error-exitcode 9 python /tmp/san3.py > gpurun_out/sanitize2_pipe_racecheck.log 2>&1; echo "pipe racecheck rc=$?"; tail -4 gpurun_out/sanitize2_pipe_racecheck.log
