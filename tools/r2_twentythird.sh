# Round 2, twenty-third GPU call (1 GPU): the final tree — full GPU suite, smoke, bench, three-pass fft2 plan re-measured with the in-place tiles.
python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python tools/gpu_bench.py fft2 2>&1 | tail -2 | cut -c1-200
SFC_FFT2_TILE2D=1 python tools/gpu_bench.py fft2 2>&1 | tail -2 | cut -c1-200
python bench.py > gpurun_out/r2y_bench_n1.json 2> gpurun_out/r2y_bench_n1.err; tail -2 gpurun_out/r2y_bench_n1.err; cut -c1-300 gpurun_out/r2y_bench_n1.json
python - <<'P'
import json
d=json.load(open('gpurun_out/r2y_bench_n1.json'))
print('value', d['value'], 'frac', d['roofline']['frac'], 'burst', d['roofline']['burst']['frac'], 'e2e', d['e2e']['value'], d['e2e'].get('pcie_frac'), d['clocks'])
for k,v in d['roofline']['others'].items(): print(k, v.get('ms', v.get('device_us')), v.get('frac'))
P
