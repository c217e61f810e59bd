"""Bank-conflict simulator for the Stockham tile kernel's shared-memory exchange.

Mirrors the index formulas in scirs_b200/csrc/fft_tile.cuh:
  thread (t, i); registers a[m] <-> element e = i + m*L/E
  stage (R, s): butterfly b of thread: ib = i + b*L/E, q = ib & (s-1),
                write pos_k = q + R*(ib-q) + k*s ; next stage reads e = i + m*L/E
  phys(t, e) = t*LP + swz(e),   swz(e) = e ^ (((e >> log2(R*s)) << log2(s)) & MASK)
A wavefront = group of G threads (G = 128B / elem_bytes); conflict degree =
max multiplicity of (phys mod G) within a group.
"""
import sys, math
from collections import Counter

def radix_plan(L, E):
    rs = []
    n = L
    while n > 1:
        r = min(E, n)
        rs.append(r)
        n //= r
    return rs

def plan_balanced(L, E=16):
    # mirror of the C++ constexpr plan: greedy 16s then remainder
    return radix_plan(L, E)

def lp_for(L, T, G):
    if T == 1:
        return L
    want = max(1, G // T) if T < G else 1
    lp = L
    while lp % G != want % G:
        lp += 1
    return lp

def sim(L, T, mode, elem_bytes, E=16):
    E = min(E, L)
    G = 128 // elem_bytes
    MASK = G - 1
    nthr_lane = L // E
    nthr = T * nthr_lane
    LP = lp_for(L, T, G)
    plan = plan_balanced(L, E)
    worst = 1
    s = 1
    report = []
    for st, R in enumerate(plan[:-1]):
        sh_rs = int(math.log2(R * s)); sh_s = int(math.log2(s))
        swz = lambda e: e ^ (((e >> sh_rs) << sh_s) & MASK)
        # writes: per (b,k) instruction
        wmax = 1; rmax = 1
        for b in range(E // R):
            for k in range(R):
                for g0 in range(0, nthr, G):
                    c = Counter()
                    for tid in range(g0, min(g0 + G, nthr)):
                        if mode == 'COL':
                            t = tid % T; i = tid // T
                        else:
                            i = tid % nthr_lane; t = tid // nthr_lane
                        ib = i + b * (L // E)
                        q = ib & (s - 1)
                        pos = q + R * (ib - q) + k * s
                        c[(t * LP + swz(pos)) % G] += 1
                    wmax = max(wmax, max(c.values()))
        for m in range(E):
            for g0 in range(0, nthr, G):
                c = Counter()
                for tid in range(g0, min(g0 + G, nthr)):
                    if mode == 'COL':
                        t = tid % T; i = tid // T
                    else:
                        i = tid % nthr_lane; t = tid // nthr_lane
                    e = i + m * (L // E)
                    c[(t * LP + swz(e)) % G] += 1
                rmax = max(rmax, max(c.values()))
        report.append((R, s, wmax, rmax))
        worst = max(worst, wmax, rmax)
        s *= R
    return worst, report, LP

if __name__ == '__main__':
    for eb in (16, 8):
        for L in (32, 64, 128, 256, 512, 1024, 2048, 4096, 8192):
            for T in (1, 2, 4, 8, 16, 32):
                if L * T // min(16, L) > 1024 or L*T*eb > 200*1024 or L*T//min(16,L) < 32:
                    continue
                for mode in ('ROW', 'COL'):
                    w, rep, LP = sim(L, T, mode, eb)
                    flag = '' if w == 1 else '  <-- CONFLICT'
                    print(f'eb={eb} L={L} T={T} {mode} LP={LP} worst={w} {rep}{flag}')
