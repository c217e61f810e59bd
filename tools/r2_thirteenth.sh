# Round 2, thirteenth GPU call (2 GPUs): pipelined slab exchange (column blocks of n2, axis-0 pass on a side stream) — parity, then bench.
SFC_TEST_WORLDS=2 timeout 1200 python -m pytest tests/test_gpu_multi.py -x -q -m gpu -s 2>&1 | tail -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 2>gpurun_out/r2m_bench_n2.err | tee gpurun_out/r2m_bench_n2.json | cut -c1-300
tail -3 gpurun_out/r2m_bench_n2.err
python - <<'P'
import json
d=json.load(open('gpurun_out/r2m_bench_n2.json'))
for k,v in d['roofline']['others'].items():
    if 'slab' in k: print(k, json.dumps(v))
P
