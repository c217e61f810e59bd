# Round 2, twelfth GPU call (1 GPU): power-of-three tiles with conflict-free lane bases (tools/bank_sim_r3.py).
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "power_of_three or 3pow13" 2>&1 | tail -3
python tools/ab_headline.py 32 1594323
SFC_R3_TL_729=3 python tools/ab_headline.py 32 1594323
python tools/ab_headline.py 256 1594323
python tools/ab_headline.py 4096 6561
python tools/ab_headline.py 512 59049
ncu --set full --clock-control none --import-source on -k regex:r3_tile -s 2 -c 2 -o gpurun_out/r2l_full_r3 -f python tools/ncu_one.py r3_13 2 > /dev/null 2>&1
python tools/summarize_ncu.py r2l_r3 gpurun_out/r2l_full_r3.ncu-rep > /dev/null 2>&1
cp profiles/r2l_r3_ncu_full.md profiles/r2l_r3_ncu_full.json gpurun_out/ 2>/dev/null
rm -f gpurun_out/r2l_full_r3.ncu-rep
cut -c1-700 gpurun_out/r2l_r3_ncu_full.md
