# Final capture of round 1: GPU tests, smoke, bench, ncu launch list of the bench command, full ncu of the top kernels
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r1c.json 2> gpurun_out/bench_r1c.err; tail -2 gpurun_out/bench_r1c.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r1c_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r1c_bench_under_ncu.log 2>&1
for c in c2c4096 dct2 c2c8192; do
  ncu --set full --clock-control none --import-source on -k regex:tile_fft -s 1 -c 1 -o gpurun_out/r1c_full_$c -f python tools/ncu_one.py $c 2 > /dev/null 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:tile_fft -s 3 -c 3 -o gpurun_out/r1c_full_blue1m -f python tools/ncu_one.py blue1m 2 > /dev/null 2>&1
ls -la gpurun_out | grep r1c
