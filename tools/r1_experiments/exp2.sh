for c in "$@"; do
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active,smsp__inst_executed.sum,launch__registers_per_thread,launch__occupancy_limit_shared_mem --clock-control none -k regex:tile_fft -c 12 --csv --log-file gpurun_out/ll_$c.csv python tools/ncu_one.py $c 2 > gpurun_out/ll_$c.txt 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/ll_$c.csv")) if len(r)>10]
hdr=rows[0]; ix={h:i for i,h in enumerate(hdr)}
agg={}
for r in rows[1:]:
    agg.setdefault((int(r[ix['ID']]),r[ix['Kernel Name']]),{})[r[ix['Metric Name']].split('.')[0]]=float(r[ix['Metric Value']].replace(',',''))
print("== $c")
for (i,k),m in sorted(agg.items()):
    t=m['gpu__time_duration']/1e3; b=(m['dram__bytes_read']+m['dram__bytes_write'])/1e9
    print(f"{i:2d} {k[21:75]:54s} {t:9.1f} us  dram {b:6.3f} GB -> {b/t*1e3:6.0f} GB/s  fp64 {m['sm__inst_executed_pipe_fp64']:4.1f}%  inst {m['smsp__inst_executed']/1e6:7.1f}M regs {m['launch__registers_per_thread']:.0f} occ_smem {m['launch__occupancy_limit_shared_mem']:.0f}")
PY
grep "step" gpurun_out/ll_$c.txt | head -8
done
