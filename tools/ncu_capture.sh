for c in c2c4096 rfft4096 irfft4096 c2c4096f32 c2c8192; do
  ncu --set full --clock-control none --import-source on -k regex:tile_fft -s 1 -c 1 -o gpurun_out/r1_full_$c -f python tools/ncu_one.py $c 2 > /dev/null 2>&1
done
ls -la gpurun_out/*.ncu-rep
