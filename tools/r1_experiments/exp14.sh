# L2 blocking experiment: shrink the four-step / Bluestein work area so consecutive passes meet in L2
for mb in 2048 256 128 96 64 48 32 16; do
  echo "--- SFC_WORK_MB=$mb"
  SFC_WORK_MB=$mb python tools/gpu_bench.py fft1m blue 2>&1 | cut -c1-150
done
