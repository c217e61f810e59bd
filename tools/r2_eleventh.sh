# Round 2, eleventh GPU call (1 GPU): shared-memory wavefronts of the power-of-three exchange patterns; cfg-1 latency knobs.
./tools/micro/smem_pattern
echo "=== cfg 1 latency"
python tools/lat_2p20.py 2>&1 | head -4 | cut -c1-250
SFC_COL_TL=2 python tools/lat_2p20.py 2>&1 | head -4 | cut -c1-250
SFC_COL_TL=8 python tools/lat_2p20.py 2>&1 | head -4 | cut -c1-250
SFC_PIPE_LATE=0 SFC_COL_TL=2 python tools/lat_2p20.py 2>&1 | head -4 | cut -c1-250
