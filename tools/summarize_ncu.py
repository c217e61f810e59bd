"""Summarise ncu reports into small tracked files under profiles/.
   usage: summarize_ncu.py <tag> <report.ncu-rep> [<report> ...]"""
import csv, subprocess, sys, os, json, collections, re

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'smsp__inst_executed.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'sass__inst_executed_local_loads', 'sass__inst_executed_local_stores',
        'lts__t_sector_hit_rate.pct', 'sm__cycles_elapsed.avg.per_second', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic']

def raw(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        res.append(d)
    return res, dict(zip(hdr, units))

def stalls(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    blocks = out.split('"Kernel Name"')
    res = []
    for b in blocks[1:]:
        rows = list(csv.reader(('"Kernel Name"' + b).splitlines()))
        hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
        st = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
        tot = collections.Counter(); ops = collections.Counter(); n = 0
        for r in rows[2:]:
            if len(r) <= ix['# Samples']: continue
            for s_ in st: tot[s_] += int(r[ix[s_]] or 0)
            ex = int(r[ix['Instructions Executed']] or 0)
            src = re.sub(r'^@!?U?P\w+\s+', '', r[ix['Source']].strip())
            op = src.split()[0].split('.')[0] if src else '?'
            ops[op] += ex; n += ex
        ssum = sum(tot.values()) or 1
        top = {k: v for k, v in ops.most_common(12)}
        for k in ("UTMALDG", "UBLKCP", "SYNCS", "UTMASTG"):  # TMA / mbarrier evidence, however few of them execute
            if ops.get(k):
                top[k] = ops[k]
        res.append(({k: round(v / ssum, 3) for k, v in tot.most_common(8)}, top, n))
    return res

tag = sys.argv[1]
summary = {}
lines = [f"# ncu --set full summaries, {tag} (clock-control none; per-launch, serialised, cold cache)\n"]
for rep in sys.argv[2:]:
    name = os.path.basename(rep).replace('.ncu-rep', '')
    rows, units = raw(rep)
    st = stalls(rep)
    if rows and len(st) > len(rows) and len(st) % len(rows) == 0:  # the source page lists every kernel once per view (SASS, PTX, ...)
        st = st[::len(st) // len(rows)]
    for i, d in enumerate(rows):
        k = d['Kernel Name']
        m = {key: d.get(key) for key in KEYS if key in d}
        t_us = float(m['gpu__time_duration.sum'].replace(',', '')) * {'ns': 1e-3, 'us': 1, 'ms': 1e3, 'msecond': 1e3, 'usecond': 1, 'nsecond': 1e-3}.get(units['gpu__time_duration.sum'], 1)
        def tobytes(key):
            v = float(m[key].replace(',', '')); u = units[key]
            return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)
        rd, wr = tobytes('dram__bytes_read.sum'), tobytes('dram__bytes_write.sum')
        lines.append(f"## {name} launch {i}: `{k}`")
        lines.append(f"- duration {t_us:.1f} us; DRAM read {rd/1e9:.3f} GB + write {wr/1e9:.3f} GB = {(rd+wr)/1e9:.3f} GB -> {(rd+wr)/t_us/1e3:.0f} GB/s under ncu")
        lines.append(f"- dram throughput {m.get('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')} % of ncu peak; FP64 pipe {m.get('sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active')} %; issue active {m.get('smsp__issue_active.avg.pct_of_peak_sustained_active')} %; warps active {m.get('sm__warps_active.avg.pct_of_peak_sustained_active')} %")
        lines.append(f"- grid {m.get('launch__grid_size')} x {m.get('launch__block_size')} thr, {m.get('launch__registers_per_thread')} regs, dyn smem {m.get('launch__shared_mem_per_block_dynamic')} {units.get('launch__shared_mem_per_block_dynamic','')}; CTAs/SM limit regs {m.get('launch__occupancy_limit_registers')} smem {m.get('launch__occupancy_limit_shared_mem')}")
        lines.append(f"- smem wavefronts {m.get('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum')}, bank conflicts {m.get('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum')}; local loads {m.get('sass__inst_executed_local_loads')} stores {m.get('sass__inst_executed_local_stores')}; L2 hit {m.get('lts__t_sector_hit_rate.pct')} %")
        if i < len(st):
            lines.append(f"- warp-instructions executed {st[i][2]}; top opcodes {st[i][1]}")
            lines.append(f"- stall sample shares {st[i][0]}")
        lines.append("")
        summary[f"{name}#{i}"] = {"kernel": k, "duration_us": round(t_us, 2), "dram_bytes_per_launch": int(rd + wr),
                                  "dram_read": int(rd), "dram_write": int(wr)}
open(f'profiles/{tag}_ncu_full.md', 'w').write('\n'.join(lines))
json.dump(summary, open(f'profiles/{tag}_ncu_full.json', 'w'), indent=1)
print('\n'.join(lines)[:6000])
