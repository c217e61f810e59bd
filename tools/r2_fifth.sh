# Round 2, fifth GPU call (8 GPUs): the multi-GPU tests at world size 8 and the bench line at N = 8.
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
nvidia-smi topo -m | head -12
SFC_TEST_WORLDS=8 timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu -s 2>&1 | tail -25
echo "=== bench n8"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 2>gpurun_out/r2e_bench_n8.err | tee gpurun_out/r2e_bench_n8.json | cut -c1-1500
tail -4 gpurun_out/r2e_bench_n8.err
