# Round 2, ninth GPU call (2 GPUs): re-validate the multi-GPU paths after the scatter-tile change, bench at N = 2, L2 panel probe.
SFC_TEST_WORLDS=2 timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu -s 2>&1 | tail -6
python tools/ab_headline.py 32768 8192
python tools/l2_panel_probe.py
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 2>gpurun_out/r2i_bench_n2.err | tee gpurun_out/r2i_bench_n2.json | cut -c1-400
tail -3 gpurun_out/r2i_bench_n2.err
