# Round 2, tenth GPU call (1 GPU): ncu of the retuned power-of-three passes and of the prime-length Bluestein passes; per-launch times.
echo "=== launch list r3_13 (32 rows)"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:tile -s 4 -c 4 --csv --log-file gpurun_out/r2j_r3_launches.csv python tools/ncu_one.py r3_13 3 > /dev/null 2>&1
cut -d, -f5,12- gpurun_out/r2j_r3_launches.csv | tail -5 | cut -c1-200
echo "=== launch list blue1m (16 rows)"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:tile -c 24 --csv --log-file gpurun_out/r2j_blue_launches.csv python tools/ncu_one.py blue1m 2 > /dev/null 2>&1
cut -d, -f5,12- gpurun_out/r2j_blue_launches.csv | tail -13 | cut -c1-200
echo "=== ncu full r3_13"
ncu --set full --clock-control none --import-source on -k regex:r3_tile -s 2 -c 2 -o gpurun_out/r2j_full_r3 -f python tools/ncu_one.py r3_13 2 > /dev/null 2>&1
python tools/summarize_ncu.py r2j_r3 gpurun_out/r2j_full_r3.ncu-rep > /dev/null 2>&1
cp profiles/r2j_r3_ncu_full.md profiles/r2j_r3_ncu_full.json gpurun_out/ 2>/dev/null
echo "=== ncu full blue1m"
ncu --set full --clock-control none --import-source on -k regex:tile -s 6 -c 6 -o gpurun_out/r2j_full_blue -f python tools/ncu_one.py blue1m 2 > /dev/null 2>&1
python tools/summarize_ncu.py r2j_blue gpurun_out/r2j_full_blue.ncu-rep > /dev/null 2>&1
cp profiles/r2j_blue_ncu_full.md profiles/r2j_blue_ncu_full.json gpurun_out/ 2>/dev/null
python tools/ncu_one.py blue1m 1 | tail -20 | cut -c1-300
ls -la gpurun_out/*.ncu-rep
