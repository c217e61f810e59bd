# Round 2, twenty-first GPU call (1 GPU): one named-barrier group per lane for the 512 x 4 / 1024 x 2 row tiles (SFC_ROW_LANE_GROUPS=1).
for n in 512 1024; do
  rows=$((268435456 / n))
  python tools/ab_headline.py $rows $n
  SFC_ROW_LANE_GROUPS=1 python tools/ab_headline.py $rows $n
done
SFC_ROW_LANE_GROUPS=1 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "golden or rows or lengths or fftn" 2>&1 | tail -2
SFC_ROW_LANE_GROUPS=1 python tools/gpu_bench.py fftn 2>&1 | grep "fftn 512" | cut -c1-160
python tools/gpu_bench.py fftn 2>&1 | grep "fftn 512" | cut -c1-160
