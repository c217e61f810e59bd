"""End-to-end timing of the scirs2-signal welch / stft callers (host buffers in, host results out) against the
oracle's per-segment loop on a bounded sample.  Writes gpurun_out/signal_bench.json.  Not a bench.py metric:
it documents what batching the reference's segment loop buys (DESIGN f-4b)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import scirs_b200.signal as sg
from oracle import signal_oracle as so

rng = np.random.default_rng(0)
res = {}
for name, n, nperseg, nover in (("welch_2p24_4096", 1 << 24, 4096, 2048), ("welch_2p24_256", 1 << 24, 256, 128),
                                ("stft_2p22_1024", 1 << 22, 1024, 512)):
    x = rng.standard_normal(n)
    fn = (lambda: sg.welch(x, 1.0, "hann", nperseg, nover)) if name.startswith("welch") else \
         (lambda: sg.stft(x, 1.0, "hann", nperseg, nover))
    fn(); fn()
    ts = []
    for _ in range(5):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    t = sorted(ts)[2]
    step = nperseg - nover
    segs = (n - nover) // step
    # CPU: the oracle's literal loop on a bounded prefix (about 2 s), scaled per segment
    m = min(segs, 400 if nperseg >= 1024 else 4000)
    xp = x[: (m - 1) * step + nperseg]
    ofn = (lambda: so.welch(xp, 1.0, "hann", nperseg, nover)) if name.startswith("welch") else \
          (lambda: so.stft(xp, 1.0, "hann", nperseg, nover, None, None, "none", False))
    t0 = time.perf_counter(); ofn(); tc = time.perf_counter() - t0
    res[name] = {"samples": n, "nperseg": nperseg, "segments": int(segs), "gpu_e2e_ms": t * 1e3,
                 "gpu_segments_per_s": segs / t, "gpu_input_GBps": n * 8 / t / 1e9,
                 "cpu_port_segments_per_s": m / tc, "cpu_sample_segments": int(m), "cpu_cores": 1}
    print(name, json.dumps(res[name]), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/signal_bench.json", "w"), indent=1)
