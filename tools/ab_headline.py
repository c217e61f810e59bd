"""Sustained (power-capped) timing of one plan: 1.5 s of the same kernel, then >= 1 s timed, clocks sampled meanwhile.
   python tools/ab_headline.py [rows n kind prec]      (knobs come from the environment: SFC_FORCE_E, SFC_PIPE, SFC_LIB_PATH, ...)"""
import os, subprocess, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scirs_b200 import FftPlan
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
kind = sys.argv[3] if len(sys.argv) > 3 else "c2c"
prec = sys.argv[4] if len(sys.argv) > 4 else "f64"
HBM = 6553.9
dev = torch.device("cuda:0"); s = torch.cuda.current_stream()
rt = torch.float64 if prec == "f64" else torch.float32
half = n // 2 + 1
n_in = {"c2c": 2 * rows * n, "r2c": rows * n, "c2r": 2 * rows * half}[kind]
n_out = {"c2c": 2 * rows * n, "r2c": 2 * rows * half, "c2r": rows * n}[kind]
x = torch.randn(n_in, device=dev, dtype=rt); y = torch.empty(n_out, device=dev, dtype=rt)
p = FftPlan([rows, n], [1], kind, prec, kind != "c2r", 1.0)
run = lambda: p.execute_device(x, y, s.cuda_stream)
lines = []
proc = subprocess.Popen(["nvidia-smi", "--id=0", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
threading.Thread(target=lambda: [lines.append(l) for l in proc.stdout], daemon=True).start()
t_end = time.time() + float(os.environ.get("AB_PRELOAD_S", "1.5"))
while time.time() < t_end:
    for _ in range(16): run()
    torch.cuda.synchronize()
n0 = len(lines)
reps = 700 if n * rows >= (1 << 27) else 2000
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(s)
for _ in range(reps): run()
e1.record(s); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
time.sleep(0.05); proc.terminate()
clk = sorted(float(l.split(",")[0]) for l in lines[n0:] if "," in l) or [0.0]
pw = sorted(float(l.split(",")[1]) for l in lines[n0:] if "," in l) or [0.0]
alg = p.info["algorithmic_bytes"]
knobs = " ".join(f"{k}={v}" for k, v in sorted(os.environ.items()) if k.startswith("SFC_"))
print(f"{kind} {prec} {rows}x{n}: {ms:.4f} ms  {alg/ms/1e6:7.1f} GB/s  {alg/ms/1e6/HBM:6.1%}  sm {clk[len(clk)//2]:.0f} MHz  {pw[len(pw)//2]:.0f} W | {knobs} | {p.describe().splitlines()[1].strip()[:110]}", flush=True)
