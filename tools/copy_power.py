"""Sustained device-to-device copy: bandwidth, SM clock and power under the same 1000 W cap the FFT kernels hit
(what the HBM roofline looks like when the measurement is not a burst).  torch's copy kernel, nothing of ours."""
import subprocess, sys, threading, time
import torch
dev = torch.device("cuda:0")
n = 1 << 29  # 4 GiB of f64 in, 4 GiB out: same footprint as the headline workload
a = torch.randn(n, device=dev, dtype=torch.float64); b = torch.empty_like(a)
lines = []
proc = subprocess.Popen(["nvidia-smi", "--id=0", "--query-gpu=clocks.sm,power.draw,clocks.mem", "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
threading.Thread(target=lambda: [lines.append(l) for l in proc.stdout], daemon=True).start()
for label, secs in (("burst (cold)", 0.0), ("sustained", 2.0)):
    t_end = time.time() + secs
    while time.time() < t_end:
        for _ in range(16): b.copy_(a)
        torch.cuda.synchronize()
    n0 = len(lines)
    reps = 10 if secs == 0 else 400
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): b.copy_(a)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    time.sleep(0.12)
    rows = [l.split(",") for l in lines[n0:] if l.count(",") >= 2]
    clk = sorted(float(r[0]) for r in rows) or [0]; pw = sorted(float(r[1]) for r in rows) or [0]; mem = sorted(float(r[2]) for r in rows) or [0]
    print(f"copy 4 GiB -> 4 GiB {label}: {ms:.4f} ms  {2 * 8 * n / ms / 1e6:.1f} GB/s  sm {clk[len(clk)//2]:.0f} MHz  mem {mem[len(mem)//2]:.0f} MHz  {pw[len(pw)//2]:.0f} W", flush=True)
proc.terminate()
