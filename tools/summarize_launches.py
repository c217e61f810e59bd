"""ncu launch list (csv from `--metrics gpu__time_duration.sum`) -> small markdown table under profiles/.
   usage: summarize_launches.py <launches.csv> <out.md> <title>"""
import csv, sys, collections, re
rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 5]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit"); mi = hdr.index("Metric Name")
def short(k):
    k = re.sub(r"void at::.*", "torch RNG / fill kernel (input generation, outside the timed region)", k)
    return k.replace("(sfc::PassParams)", "").replace("void ", "")[:110]
tot = collections.OrderedDict(); order = []
for r in rows[1:]:
    if r[mi] != "gpu__time_duration.sum": continue
    v = float(r[vi].replace(",", "")); v = v / 1000.0 if r[ui] in ("ns", "nsecond") else v
    k = short(r[ki]); order.append((k, v))
    a = tot.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
total = sum(a[1] for a in tot.values())
with open(sys.argv[2], "w") as f:
    f.write(f"# {sys.argv[3]}\n\n`ncu --metrics gpu__time_duration.sum --clock-control none` — per-launch times are cold-cache and serialised; compare SHARES.\n\n")
    f.write(f"total device time in the capture: {total/1000:.2f} ms over {len(order)} launches\n\n| kernel | launches | total ms | share | mean us |\n|---|---:|---:|---:|---:|\n")
    for k, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        f.write(f"| `{k}` | {n} | {t/1000:.3f} | {t/total:.1%} | {t/n:.1f} |\n")
    f.write("\n## first 30 launches in order\n\n| id | kernel | us |\n|---:|---|---:|\n")
    for i, (k, v) in enumerate(order[:30]):
        f.write(f"| {i} | `{k}` | {v:.1f} |\n")
print("wrote", sys.argv[2])
