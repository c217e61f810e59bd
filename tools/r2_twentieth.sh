# Round 2, twentieth GPU call (8 GPUs): bench line at N = 8 with the round's final kernels (weak-scaling headline, slab fftn 512^3 / 1024^3
# with parity), then the device-buffer multi-GPU test at world size 8.
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --steps 20 --warmup 5 2>gpurun_out/r2v_bench_n8.err | tee gpurun_out/r2v_bench_n8.json | cut -c1-300
tail -2 gpurun_out/r2v_bench_n8.err
python - <<'P'
import json
d=json.load(open('gpurun_out/r2v_bench_n8.json'))
print('value', d['value'], 'frac', d['roofline']['frac'], 'e2e', d['e2e'])
for k,v in d['roofline']['others'].items():
    if 'slab' in k: print(k, json.dumps(v)[:420])
P
SFC_TEST_WORLDS=8 timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -m gpu -s -k "device_buffers" 2>&1 | tail -4
