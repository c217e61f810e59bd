"""Bank-conflict check of the PADDED exchange layout (mirrors fft_tile.cuh):
   phys(t, e) = t*LP + e + (e >> log2(R*S)) * (S if S < G else 0)
"""
import math
from collections import Counter

def sim(L, T, mode, eb, E=16):
    E = min(E, L); G = 128 // eb
    TPL = L // E; NT = T * TPL
    plan = []; n = L
    while n > 1:
        r = min(E, n); plan.append(r); n //= r
    # lane pitch: padded length, then COL rule
    padded = L + (L // plan[0] if len(plan) > 1 else 0) + 1
    want = (G // T) % G if T < G else 1
    LP = padded
    while LP % G != want: LP += 1
    worst = 1; rep = []; s = 1
    for R in plan[:-1]:
        sh = int(math.log2(R * s)); padw = s if s < G else 0
        ph = lambda e: e + (e >> sh) * padw
        wmax = rmax = 1
        def tmap(tid):
            return (tid % T, tid // T) if mode == 'COL' else (tid // TPL, tid % TPL)
        for b in range(E // R):
            for k in range(R):
                for g0 in range(0, NT, G):
                    c = Counter()
                    for tid in range(g0, min(g0 + G, NT)):
                        t, i = tmap(tid); ib = i + b * TPL; q = ib & (s - 1)
                        c[(t * LP + ph(q + R * (ib - q) + k * s)) % G] += 1
                    wmax = max(wmax, max(c.values()))
        for m in range(E):
            for g0 in range(0, NT, G):
                c = Counter()
                for tid in range(g0, min(g0 + G, NT)):
                    t, i = tmap(tid)
                    c[(t * LP + ph(i + m * TPL)) % G] += 1
                rmax = max(rmax, max(c.values()))
        rep.append((R, s, wmax, rmax)); worst = max(worst, wmax, rmax); s *= R
    return worst, rep, LP

if __name__ == '__main__':
    for eb in (16, 8):
        for L in (32, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384):
            for T in (1, 2, 4, 8, 16, 32, 64, 128):
                NT = L * T // min(16, L)
                if NT > 1024 or NT < 128 or L * T * eb > 140 * 1024: continue
                for mode in ('ROW', 'COL'):
                    w, rep, LP = sim(L, T, mode, eb)
                    print(f'eb={eb} L={L} T={T} {mode} LP={LP} worst={w} {rep}' + ('' if w == 1 else '  <-- CONFLICT'))
