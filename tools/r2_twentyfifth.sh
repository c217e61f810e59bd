# Round 2, twenty-fifth GPU call (1 GPU): mirrored r2c with the in-place middle stage, three alternating A/B pairs (build_ab = -DSFC_INPLACE_MID=2).
for i in 1 2 3; do
python tools/ab_headline.py 65536 4096 r2c
SFC_LIB_PATH=$PWD/build_ab/libscirs2_fft_cuda.so python tools/ab_headline.py 65536 4096 r2c
done
