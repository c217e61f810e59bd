# Round 2, nineteenth GPU call (2 GPUs): the multi-GPU paths with the round's final kernels (in-place tiles, windows) + bench at N = 2.
SFC_TEST_WORLDS=2 timeout 1200 python -m pytest tests/test_gpu_multi.py -x -q -m gpu -s 2>&1 | tail -6
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 20 --warmup 5 2>gpurun_out/r2u_bench_n2.err | tee gpurun_out/r2u_bench_n2.json | cut -c1-300
tail -2 gpurun_out/r2u_bench_n2.err
python - <<'P'
import json
d=json.load(open('gpurun_out/r2u_bench_n2.json'))
print('frac', d['roofline']['frac'], 'e2e', d['e2e']['value'])
for k,v in d['roofline']['others'].items():
    if 'slab' in k: print(k, json.dumps(v)[:400])
P
