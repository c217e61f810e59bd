# Round 2, second GPU call (2 GPUs): multi-GPU tests first, then the full suite, bench at N=1 and N=2, A/B of the headline kernel.
nvidia-smi --query-gpu=index,name,clocks.sm,power.limit --format=csv
nvidia-smi topo -m | head -20
echo "=== multi"
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_threads.py -x -q -m gpu -s 2>&1 | tail -40
echo "=== full suite"
SFC_TEST_EXPERIMENTAL=1 timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
echo "=== bench n1"
timeout 600 python bench.py 2>gpurun_out/r2b_bench_n1.err | tee gpurun_out/r2b_bench_n1.json | cut -c1-3000
echo "=== bench n2"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 2>gpurun_out/r2b_bench_n2.err | tee gpurun_out/r2b_bench_n2.json | cut -c1-3000
tail -5 gpurun_out/r2b_bench_n2.err
echo "=== A/B headline"
for v in "SFC_FORCE_E=0" "SFC_FORCE_E=8" "SFC_ROW_SPLIT=4096"; do
  echo "--- $v"
  env $v timeout 300 python tools/gpu_bench.py c2c4096 sizes 2>&1 | grep -E "65536x4096 f64|131072x2048"
done
