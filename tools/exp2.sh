for c in fft1m64 fft2_8192 c2c8192; do
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:tile_fft -c 8 --csv --log-file gpurun_out/ll_$c.csv python tools/ncu_one.py $c 2 > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/ll_$c.csv")) if len(r)>10]
hdr=rows[0]; ix={h:i for i,h in enumerate(hdr)}
agg={}
for r in rows[1:]:
    agg.setdefault((r[ix['ID']],r[ix['Kernel Name']][:60]),{})[r[ix['Metric Name']]]=r[ix['Metric Value']]
print("== $c")
for (i,k),m in agg.items():
    print(i,k,{a.split('.')[0][-28:]:b for a,b in m.items()})
PY
done
