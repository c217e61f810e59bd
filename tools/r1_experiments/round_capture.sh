# Round capture: GPU tests, smoke, bench, ncu launch list of the bench command, full ncu of the top kernels.
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()"
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; tail -2 gpurun_out/bench_r1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r1_bench_under_ncu.log 2>&1
for c in c2c4096 rfft4096 irfft4096 fftn512 c2c4096f32; do
  ncu --set full --clock-control none --import-source on -k regex:tile_fft -s 3 -c 3 -o gpurun_out/r1_full_$c -f python tools/ncu_one.py $c 2 > /dev/null 2>&1
done
ls -la gpurun_out | tail -12
