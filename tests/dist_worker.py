"""One rank of a multi-GPU check, started N times by tests/test_gpu_multi.py (or by hand) — NOT through torchrun: the
library does its own rendezvous (sfc_comm_init_rank over POSIX shared memory), so the `host` cases import neither
torch nor torch.distributed.

    python tests/dist_worker.py --rank R --world P --name JOB --case host|device [--sizes 64,256]

Prints one JSON line per check: {"rank":, "case":, "shape":, "layout":, "dir":, "prec":, "rel_l2":, "ok":}.
Exit code 0 iff every check of this rank passed."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scirs_b200 as sb  # noqa: E402
from scirs_b200 import _lib  # noqa: E402
from scirs_b200.distributed import Communicator, DistPlan  # noqa: E402


def rel(a, b):
    return float(np.linalg.norm((a - b).ravel()) / np.linalg.norm(b.ravel()))


def volume(shape, seed, prec):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal(shape) + 1j * rng.standard_normal(shape)
    return x.astype(np.complex64) if prec == "f32" else x


def reference(x, forward, prec):
    """Small volumes: the oracle (scirs2-fft semantics restated on the CPU).  Larger: the single-GPU plan of this
    library, which tests/test_gpu_parity.py pins against the oracle at 512^3 / 1024^3."""
    x64 = x.astype(np.complex128)
    if x.size <= 1 << 21:
        from oracle import scirs2_fft_oracle as orc

        return orc.fftn(x64) if forward else orc.ifftn(x64, None, None, "forward")  # "forward" on ifftn = unscaled
    p = sb.FftPlan(list(x.shape), [0, 1, 2], "c2c", "f64", forward)
    return p.execute(x64).reshape(x.shape)


def emit(ok_all, **kw):
    print(json.dumps(kw), flush=True)
    return ok_all and kw["ok"]


def case_host(comm, rank, P, sizes):
    """sfc_dist_exec_host in rank mode: this rank's slab from / to host memory; nothing but numpy + the C ABI."""
    ok = True
    shapes = [(n, n, n) for n in sizes] + [(16 * P, 8 * P, 64), (8 * P, 32 * P, 32)]
    for shape in shapes:
        for prec in ("f64", "f32"):
            tol = 1e-12 if prec == "f64" else 1e-5
            x = volume(shape, 7 + shape[0], prec)
            for forward in (True, False):
                ref = None
                for layout in ("transposed", "natural"):
                    plan = DistPlan(comm, shape, [0, 1, 2], "slab", layout, "c2c", prec, forward, 1.0)
                    s0, s1 = shape[0] // P, shape[1] // P
                    assert plan.local_in_shape == (s0, shape[1], shape[2])
                    mine = np.ascontiguousarray(x[rank * s0:(rank + 1) * s0])
                    out = np.empty(plan.local_out_shape, dtype=x.dtype)
                    worst = 0.0
                    for it in range(3):  # consecutive calls exercise both receive buffers and the epoch flags
                        out[...] = 0
                        plan.execute_host(mine, out)
                        if ref is None:
                            ref = reference(x, forward, prec)
                        want = ref[:, rank * s1:(rank + 1) * s1, :] if layout == "transposed" else ref[rank * s0:(rank + 1) * s0]
                        worst = max(worst, rel(out.astype(np.complex128), want))
                    info = plan.info
                    assert info["world"] == P and info["num_exchanges"] == (0 if P == 1 else (1 if layout == "transposed" else 2))
                    ok = emit(ok, rank=rank, case="host", shape=list(shape), layout=layout, dir="fwd" if forward else "inv",
                              prec=prec, rel_l2=worst, ok=bool(worst <= tol))
                    plan.close()
    # batch split: every rank transforms its own rows, nothing is exchanged
    B, n = 6 * P + 1, 1000  # ragged split (last rank short) and a Bluestein length
    rng = np.random.default_rng(11)
    sig = rng.standard_normal((B, n)) + 1j * rng.standard_normal((B, n))
    plan = DistPlan(comm, [B, n], [1], "batch_split")
    per = -(-B // P)
    lo, hi = min(rank * per, B), min((rank + 1) * per, B)
    assert plan.local_in_shape == (hi - lo, n), (plan.local_in_shape, hi - lo)
    out = np.empty((hi - lo, n), dtype=np.complex128)
    if hi > lo:
        plan.execute_host(np.ascontiguousarray(sig[lo:hi]), out)
        e = rel(out, np.fft.fft(sig[lo:hi], axis=1))
    else:
        e = 0.0
    ok = emit(ok, rank=rank, case="host", shape=[B, n], layout="batch_split", dir="fwd", prec="f64", rel_l2=e, ok=bool(e <= 1e-12))
    plan.close()
    return ok


def case_device(comm, rank, P, sizes):
    """sfc_dist_exec_device on the caller's stream with torch tensors as device memory (no torch.distributed)."""
    import torch

    torch.cuda.set_device(rank)
    ok = True
    for n in sizes:
        shape = (n, n, n)
        s0 = s1 = n // P
        g = torch.Generator(device="cuda").manual_seed(99)
        full = torch.view_as_complex(torch.randn(n, n, n, 2, dtype=torch.float64, device="cuda", generator=g))
        ref = torch.empty_like(full)
        sb.FftPlan([n, n, n], [0, 1, 2], "c2c", "f64").execute_device(full, ref, torch.cuda.current_stream().cuda_stream)
        mine = full[rank * s0:(rank + 1) * s0].contiguous()
        side = torch.cuda.Stream()
        # chunks: 0 = library default (one scatter + one wait), 4 / 2 / 8 = pipelined exchange over that many column
        # blocks of n2 (8 = the most the flag page holds)
        for layout, chunks in (("transposed", 0), ("natural", 0), ("transposed", 4), ("natural", 4), ("transposed", 8), ("natural", 2)):
            plan = DistPlan(comm, shape, [0, 1, 2], "slab", layout, chunks=chunks)
            assert 1 <= plan.info["chunks"] <= max(chunks, 1), plan.info
            if P > 1 and n >= 512 and chunks:
                assert plan.info["chunks"] == chunks, plan.info  # BASELINE configs[4] sizes take every block count
            want = ref[:, rank * s1:(rank + 1) * s1, :] if layout == "transposed" else ref[rank * s0:(rank + 1) * s0]
            outs = {"plain": torch.empty(plan.local_out_shape, dtype=torch.complex128, device="cuda")}
            sym = None
            if layout == "natural":
                # a symmetric allocation as output: the second exchange stores straight into it (no window, no copy)
                sym = comm.alloc(mine.numel() * 16)
                outs["symmetric"] = sym
            for kind, out in outs.items():
                worst = 0.0
                for it, st in enumerate((torch.cuda.current_stream(), side, side)):
                    st.wait_stream(torch.cuda.current_stream())
                    plan.execute_device(mine, out, st.cuda_stream)
                    st.synchronize()
                    if kind == "symmetric":
                        got = torch.empty(plan.local_out_shape, dtype=torch.complex128, device="cuda")
                        from scirs_b200.distributed import _as_tensor

                        got.view(torch.float64).reshape(-1).copy_(_as_tensor(torch, out, mine.numel() * 2, torch.float64))
                    else:
                        got = out
                    worst = max(worst, float((got - want).norm() / want.norm()))
                ok = emit(ok, rank=rank, case="device", shape=list(shape), layout=layout + ":" + kind, dir="fwd", prec="f64",
                          chunks=int(plan.info["chunks"]), rel_l2=worst, ok=bool(worst <= 1e-12))
            comm.barrier()
            if sym is not None:
                comm.free(sym)
            plan.close()
        del full, ref
    return ok


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rank", type=int, required=True)
    ap.add_argument("--world", type=int, required=True)
    ap.add_argument("--name", required=True)
    ap.add_argument("--case", default="host")
    ap.add_argument("--sizes", default="64")
    a = ap.parse_args()
    sizes = [int(v) for v in a.sizes.split(",") if v]
    lib = _lib.load()
    sb.error.check(lib.sfc_init(a.rank))
    comm = Communicator.rank_mode(a.name, a.rank, a.world, a.rank)
    assert comm.size() == a.world and comm.rank() == a.rank
    ok = case_host(comm, a.rank, a.world, sizes) if a.case == "host" else case_device(comm, a.rank, a.world, sizes)
    comm.barrier()
    comm.close()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
