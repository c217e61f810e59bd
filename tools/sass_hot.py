"""Where do a kernel's warp-stall samples fall?  Reads `ncu -i X.ncu-rep --page source --csv --print-source sass`."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr)]
S = ci['# Samples']
num = lambda v: int(v) if v.strip().isdigit() else 0
tot = sum(num(r[S]) for r in data)
print('kernel', rows[0][1][:100]); print('total samples', tot, 'instructions', len(data))
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
for r in sorted(data, key=lambda r: -num(r[S]))[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    st = sorted(((num(r[ci[h]]), h[6:]) for h in stalls), reverse=True)[:2]
    print(f"{data.index(r):5d} {num(r[S]):6d} {num(r[S])/tot:6.1%}  {st[0][1]:10s} {r[ci['Source']][:80]}")
step = 200
for k in range(0, len(data), step):
    blk = data[k:k + step]
    s = sum(num(r[S]) for r in blk)
    ops = {}
    for r in blk:
        w = r[ci['Source']].split()
        op = (w[1] if w and w[0].startswith('@') and len(w) > 1 else (w[0] if w else '')).split('.')[0]
        ops[op] = ops.get(op, 0) + 1
    t = sorted(ops.items(), key=lambda x: -x[1])[:4]
    print(f"[{k:5d}] {s:6d} {s/tot:6.1%} ", t)
