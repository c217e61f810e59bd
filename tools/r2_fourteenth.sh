# Round 2, fourteenth GPU call (2 GPUs): pipelined slab exchange, side stream at the highest priority vs. default; 1 / 2 / 4 / 8 blocks.
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 tools/slab_ab.py 512 1024 2>&1 | grep -E "slab fftn|Error|error" | cut -c1-400
SFC_SLAB_SIDE_PRIORITY=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 tools/slab_ab.py 512 2>&1 | grep -E "slab fftn|Error|error" | cut -c1-400
