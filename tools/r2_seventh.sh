# Round 2, seventh GPU call (1 GPU): power-of-three tile widths, cfg-1 latency, UTMALDG evidence, tuner + C++ mirror tests.
echo "=== r3 tile widths, 3^13 x 32 (sustained)"
python tools/ab_headline.py 32 1594323
for a in 6 4 3; do SFC_R3_TL_729=$a python tools/ab_headline.py 32 1594323; done
for b in 2 1; do SFC_R3_TL_2187=$b python tools/ab_headline.py 32 1594323; done
SFC_R3_TL_729=4 SFC_R3_TL_2187=2 python tools/ab_headline.py 32 1594323
python tools/ab_headline.py 65536 2187
SFC_R3_TL_2187=1 python tools/ab_headline.py 65536 2187
SFC_R3_TL_2187=2 python tools/ab_headline.py 65536 2187
python tools/ab_headline.py 196608 729
SFC_R3_TL_729=3 python tools/ab_headline.py 196608 729
SFC_R3_TL_729=4 python tools/ab_headline.py 196608 729
echo "=== cfg 1 latency"
python tools/lat_2p20.py | head -1
SFC_COL_SMEM_KB=40 python tools/lat_2p20.py | head -4
SFC_PIPE_LATE=0 python tools/lat_2p20.py | head -1
echo "=== tests"
timeout 600 python -m pytest tests/test_auto_tuning.py tests/test_cpp_mirror.py tests/test_planning_adaptive.py -q -m gpu 2>&1 | tail -4
echo "=== ncu: tensor-map pass of the 2^20 four-step plan"
ncu --set full --clock-control none --import-source on -k regex:tile_fft -s 2 -c 1 -o gpurun_out/r2_full_tmap -f python tools/ncu_one.py fft1m64 2 > /dev/null 2>&1
python tools/summarize_ncu.py r2tmap gpurun_out/r2_full_tmap.ncu-rep > /dev/null 2>&1
cp profiles/r2tmap_ncu_full.md profiles/r2tmap_ncu_full.json gpurun_out/ 2>/dev/null
grep -E "^## |top opcodes" gpurun_out/r2tmap_ncu_full.md | cut -c1-600
rm -f gpurun_out/r2_full_tmap.ncu-rep
