# Round capture (second half of round 1): ncu launch list of the bench command + full captures of the top kernels
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r1b_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r1b_bench_under_ncu.log 2>&1
for c in c2c4096 blue1m fft1m64 c2c8192; do
  ncu --set full --clock-control none --import-source on -k regex:tile_fft -s 3 -c 3 -o gpurun_out/r1b_full_$c -f python tools/ncu_one.py $c 2 > /dev/null 2>&1
done
ls -la gpurun_out | grep r1b
