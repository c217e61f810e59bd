# Round 2, twenty-fourth GPU call (2 GPUs): multi-GPU suite on the final tree (after the from-scratch rebuild).
SFC_TEST_WORLDS=2 timeout 1200 python -m pytest tests/test_gpu_multi.py -x -q -m gpu -s 2>&1 | tail -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps 5 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('N=2 value', d['value'], 'frac', d['roofline']['frac'])
for k,v in d['roofline']['others'].items():
    if 'slab' in k: print(k, v['ms'], v['parity_rel_l2'])
"
