"""Shared-memory wavefronts of the power-of-three tiles' exchange layout (r3_tile.cuh, R3Cfg::lane_base), f64.

Model (matches the ncu source page of the round-2 tiles, where every LDS.128 / STS.128 of a strided tile cost exactly twice
its ideal wavefronts): a warp-wide 128-bit access is served in groups of eight consecutive threads, one wavefront per group
when the eight 16-byte slots differ modulo 8.  Prints, per registered tile and thread mapping, wavefronts / ideal over every
access of every stage, for the plain pitch L (round 2, first cut) and for lane_base()."""
INPLACE = False  # the kernel default
TILES = [(9, 243), (9, 486), (27, 81), (27, 162), (81, 27), (81, 54), (243, 9), (243, 18), (729, 3), (729, 6), (2187, 1), (2187, 2), (2187, 3)]


def lane_base(L, TL, t, col, padded=True):
    if not padded:
        return t * L
    LP = (L + 7) // 8 * 8
    r = (((TL & 7) * t) if TL & 1 else (((TL // 2) & 7) * (t >> 1) + 4 * (t & 1))) if col else t * L
    return t * LP + (r & 7)


def mapping(mode, tid, TL, TPL):
    return (tid % TL, tid // TL) if mode == "col" else (tid // TPL, tid % TPL)


def stage_accesses(L):
    """yields (kind, f) with f(i) -> slot offsets inside a lane for butterfly index i (one list entry per instruction); the
    middle stage in front of the last one works in place for L >= 729 with INPLACE (r3_stages, -DSFC_R3_INPLACE_MID=1)"""
    TPL = L // 9
    stages, S = [], 1
    while True:
        R = 9 if L // S >= 9 else L // S
        if S * R == L:
            break
        stages.append((S, R))
        S *= R
    for idx, (S, R) in enumerate(stages):
        RN = L // (S * R)
        inplace = INPLACE and R == 9 and S > 1 and L >= 729 and idx == len(stages) - 1 and RN in (3, 9)
        NB = 9 // R
        if inplace:
            yield "write in place S=%d" % S, (lambda i: [i + k * TPL for k in range(9)])
            yield "read permuted", (lambda i, S=S, RN=RN: [(i % S) + TPL * (i // S) + S * ((RN * m) // 9) + TPL * ((RN * m) % 9) for m in range(9)])
            continue

        def wr(i, S=S, R=R, NB=NB):
            out = []
            for b in range(NB):
                ib = i + b * TPL
                q = ib % S
                out += [q + R * (ib - q) + k * S for k in range(R)]
            return out
        yield "write S=%d" % S, wr
        yield "read after S=%d" % S, lambda i: [i + m * TPL for m in range(9)]


def wavefronts(L, TL, wmode, rmode, layout):
    """layout: 'plain' (pitch L), 'col' / 'row' (lane_base with col = True / False)"""
    TPL = L // 9
    NT = TL * TPL
    tot = ideal = 0
    for kind, f in stage_accesses(L):
        mode = wmode if kind == "write S=1" else rmode  # only the first write uses the input mapping
        per_thread = []
        for tid in range(NT):
            t, i = mapping(mode, tid, TL, TPL)
            per_thread.append([lane_base(L, TL, t, layout == "col", layout != "plain") + o for o in f(i)])
        for inst in range(len(per_thread[0])):
            for g0 in range(0, NT, 8):
                cnt = {}
                for tid in range(g0, min(g0 + 8, NT)):
                    r = per_thread[tid][inst] & 7
                    cnt[r] = cnt.get(r, 0) + 1
                tot += max(cnt.values())
                ideal += 1
    return tot / ideal


def kernel_choice(L, wm, rm):
    """the rule in r3_tile_kernel"""
    return "col" if (wm == "col" and rm == "col") or ((wm == "col" or rm == "col") and L >= 243) else "row"


if __name__ == "__main__":
    for L, TL in TILES:
        if L == 9:
            continue  # single stage: no exchange
        for wm, rm in (("row", "row"), ("col", "col"), ("row", "col"), ("col", "row")):
            res = {lay: wavefronts(L, TL, wm, rm, lay) for lay in ("plain", "row", "col")}
            pick = kernel_choice(L, wm, rm)
            print(f"L={L:5d} TL={TL:3d} {wm}->{rm}:  plain pitch {res['plain']:5.2f}  row bases {res['row']:5.2f}  col bases {res['col']:5.2f}"
                  f"  (x ideal)   kernel uses {pick}: {res[pick]:5.2f}")
