"""Aggregate executed instructions per opcode from an ncu report's source page."""
import csv, subprocess, sys, collections, re
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
ops = collections.Counter(); tot = 0; samples = collections.Counter(); stot = 0
for r in rows[2:]:
    if len(r) <= ix['Instructions Executed']: continue
    n = int(r[ix['Instructions Executed']] or 0)
    src = re.sub(r'^@!?U?P\w+\s+', '', r[ix['Source']].strip())
    op = src.split()[0] if src else '?'
    op = '.'.join(op.split('.')[:2]) if op.startswith(('LDG', 'STG', 'LDS', 'STS', 'IMAD', 'LDL', 'STL')) else op.split('.')[0]
    ops[op] += n; tot += n
    sm = int(r[ix['# Samples']] or 0); samples[op] += sm; stot += sm
nthreads_warps = float(sys.argv[2]) if len(sys.argv) > 2 else None
print(f"total warp-instructions {tot}")
for op, n in ops.most_common(28):
    per = f" per-warp {n / nthreads_warps:8.1f}" if nthreads_warps else ""
    print(f"{op:14s} {n:12d} {n / tot:6.1%}{per}   samples {samples[op] / max(stot,1):6.1%}")
