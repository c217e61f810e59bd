# Round 2, eighteenth GPU call (1 GPU): the round's final single-GPU evidence — full GPU suite, smoke, r3 in-place A/B, bench,
# ncu launch list of the bench command, ncu --set full of the top kernels.
python -m pytest tests -x -q -m gpu 2>&1 | tail -6
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "=== power-of-three tiles, in-place middle stage (build_ab = same sources, -DSFC_INPLACE_MID=0, before the lane-base fix is NOT in it)"
python tools/ab_headline.py 256 1594323
python tools/ab_headline.py 196608 729
python tools/ab_headline.py 65536 2187
echo "=== fft2 / fftn with the in-place tiles"
python tools/gpu_bench.py fft2 fftn 2>&1 | tail -6 | cut -c1-200
echo "=== bench"
python bench.py > gpurun_out/r2t_bench_n1.json 2> gpurun_out/r2t_bench_n1.err; tail -2 gpurun_out/r2t_bench_n1.err; cut -c1-600 gpurun_out/r2t_bench_n1.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2t_bench_ref.json 2>/dev/null; cut -c1-400 gpurun_out/r2t_bench_ref.json
echo "=== ncu"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r2_bench_under_ncu.log 2>&1
for c in c2c4096 fft1m64 r3_13 rfft4096; do
  ncu --set full --clock-control none --import-source on -k regex:tile -s 1 -c 2 -o gpurun_out/r2_full_$c -f python tools/ncu_one.py $c 2 > /dev/null 2>&1
done
python tools/summarize_ncu.py r2 gpurun_out/r2_full_c2c4096.ncu-rep gpurun_out/r2_full_fft1m64.ncu-rep gpurun_out/r2_full_r3_13.ncu-rep gpurun_out/r2_full_rfft4096.ncu-rep > /dev/null 2>&1
cp profiles/r2_ncu_full.md profiles/r2_ncu_full.json gpurun_out/ 2>/dev/null
python tools/summarize_launches.py gpurun_out/r2_launches.csv gpurun_out/r2_launches_summary.md "ncu launch list of bench.py --steps 2 --warmup 3 (round 2, final)" 2>&1 | tail -1
rm -f gpurun_out/r2_full_fft1m64.ncu-rep gpurun_out/r2_full_rfft4096.ncu-rep gpurun_out/r2_full_r3_13.ncu-rep
grep -E "^## |duration|top opcodes" gpurun_out/r2_ncu_full.md | cut -c1-400 | head -12
