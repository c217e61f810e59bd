timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "pow2 or four_step or bluestein or fftn_family or fft2_family or golden" 2>&1 | tail -5
for pz in 0 1; do echo "=== SFC_PIPE=$pz"; SFC_PIPE=$pz timeout 600 python tools/gpu_bench.py c2c4096 fft1m blue fftn fft2 2>&1 | grep -v "batch 1 \|f32" | cut -c1-112; done
