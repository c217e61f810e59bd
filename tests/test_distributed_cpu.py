"""CPU, world_size 2, gloo: the host-side logic of the multi-GPU paths (partitioning, pack order,
exchange, reassembly) with a numpy local transform standing in for the CUDA library."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from scirs_b200.distributed import slab_partition, split_batch, pack_for_exchange


def test_partition_helpers():
    # distributed.rs:356-362: ceil split, last rank may be short or empty
    assert [slab_partition(10, 4, r) for r in range(4)] == [(0, 3), (3, 3), (6, 3), (9, 1)]
    assert [slab_partition(512, 8, r) for r in range(8)] == [(64 * r, 64) for r in range(8)]
    assert [split_batch(65536, 8, r) for r in range(8)] == [(8192 * r, 8192) for r in range(8)]
    assert [split_batch(5, 4, r) for r in range(4)] == [(0, 2), (2, 2), (4, 1), (5, 0)]
    y = np.arange(2 * 4 * 3).reshape(2, 4, 3).astype(np.complex128)
    p = pack_for_exchange(y, 2)
    assert p.shape == (2, 2, 2, 3)
    assert np.array_equal(p[1], y[:, 2:4, :])


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import scirs2_fft_oracle as orc
        from scirs_b200.distributed import SlabFFT3D

        rng = np.random.default_rng(6)
        full = rng.standard_normal((n, n, n)) + 1j * rng.standard_normal((n, n, n))
        start, cnt = slab_partition(n, world, rank)

        def local(a, axes):
            return orc.fftn(a, None, list(axes))

        f = SlabFFT3D(n, n, n, local_transform=local)
        out = f.forward_host(np.ascontiguousarray(full[start:start + cnt]))
        ref = orc.fftn(full)  # the reference's fftn(&a, None, None, None, None, None)
        s1 = n // world
        err = orc.rel_l2(out, ref[:, rank * s1:(rank + 1) * s1, :])
        # batch split: every rank transforms its own signals, nothing is exchanged
        b0, bc = split_batch(12, world, rank)
        sig = rng.standard_normal((12, 32))
        mine = np.stack([orc.rfft(r) for r in sig[b0:b0 + bc]])
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        err2 = orc.rel_l2(np.concatenate(gathered), np.stack([orc.rfft(r) for r in sig]))
        q.put((rank, err, err2))
    finally:
        dist.destroy_process_group()


def test_slab_fftn_two_ranks_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 16, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err, err2 in res:
        assert err < 1e-13, (rank, err)
        assert err2 == 0.0
