"""Multi-GPU check (run under torchrun on >= 2 GPUs):  slab fftn against the single-GPU plan."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scirs_b200 as sb
from scirs_b200 import _lib
from scirs_b200.distributed import SlabFFT3D, bench_slab_fftn


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    sb.error.check(_lib.load().sfc_init(local))
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    for n in (64, 256):
        g = torch.Generator(device="cuda").manual_seed(99)
        full = torch.view_as_complex(torch.randn(n, n, n, 2, dtype=torch.float64, device="cuda", generator=g))
        ref = torch.empty_like(full)
        sb.FftPlan([n, n, n], [0, 1, 2], "c2c", "f64").execute_device(full, ref, torch.cuda.current_stream().cuda_stream)
        s0, s1 = n // world, n // world
        for mode in ("p2p", "nccl"):
            f = SlabFFT3D(n, n, n, mode=mode)
            out = torch.empty(n, s1, n, dtype=torch.complex128, device="cuda")
            for it in range(2):
                f.forward_device(full[rank * s0:(rank + 1) * s0].contiguous(), out)
            torch.cuda.synchronize()
            want = ref[:, rank * s1:(rank + 1) * s1, :]
            err = float((out - want).norm() / want.norm())
            print(f"rank {rank} n={n} mode={mode} rel_l2={err:.3e}", flush=True)
            ok = ok and err < 1e-12
            f.close()
            dist.barrier()
    for n in (512, 1024) if world >= 2 else ():
        for mode in ("p2p", "nccl"):
            try:
                r = bench_slab_fftn(n, steps=5, warmup=2, mode=mode)
                if rank == 0:
                    print(r, flush=True)
            except Exception as ex:
                if rank == 0:
                    print("bench failed", n, mode, ex, flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
