"""A/B of the pipelined slab exchange (torchrun, one rank per GPU): slab fftn n^3 with 1 / 2 / 4 / 8 column blocks.
usage: torchrun --nproc-per-node P tools/slab_ab.py [n ...]   (knobs from the environment, e.g. SFC_SLAB_SIDE_PRIORITY=0)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from scirs_b200.distributed import bench_slab_fftn, Communicator

rank = int(os.environ["RANK"]); torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
import scirs_b200 as sb
from scirs_b200 import _lib
sb.error.check(_lib.load().sfc_init(int(os.environ["LOCAL_RANK"])))
comm = Communicator.from_env()
sizes = [int(v) for v in sys.argv[1:]] or [512]
knobs = " ".join(f"{k}={v}" for k, v in sorted(os.environ.items()) if k.startswith("SFC_"))
for n in sizes:
    for lay in ("transposed", "natural"):
        row = []
        for ch in (1, 2, 4, 8):
            r = bench_slab_fftn(n, steps=10, warmup=3, layout=lay, comm=comm, check_parity=(ch == 4), min_seconds=0.2, chunks=ch)
            p = r["pipelined"]
            row.append(f"{ch}: {r['ms_per_step']:.4f}" + (f" (B blocks {p['stage_ms']['fft_axis1_scatter_blocks']:.3f} tail {p['stage_ms']['fft_axis0_tail']:.3f})" if p else f" (A {r['stage_ms']['fft_axis2']:.3f} B {r['stage_ms']['fft_axis1_scatter']:.3f} C {r['stage_ms']['fft_axis0']:.3f})"))
        if rank == 0:
            print(f"slab fftn {n}^3 x{dist.get_world_size()} {lay:10s} ms by chunks | " + " | ".join(row) + " | " + knobs, flush=True)
comm.close()
dist.destroy_process_group()
