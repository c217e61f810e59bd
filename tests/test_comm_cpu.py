"""CPU: the library's own rendezvous (sfc_comm_init_rank over POSIX shared memory — no Python, torch or NCCL in it)
driven by several processes with host-only communicators (device = -1): all-gather, barrier, error behaviour.
Mirrors the contract of `trait Communicator` (scirs2-fft/src/distributed.rs:85-103: barrier / size / rank)."""
import ctypes as C
import multiprocessing as mp
import os
import time

import numpy as np
import pytest


def _worker(rank, world, name, q, nbytes):
    try:
        from scirs_b200 import _lib

        lib = _lib.load()
        comm = C.c_void_p()
        rc = lib.sfc_comm_init_rank(C.byref(comm), name.encode(), rank, world, -1)
        if rc != 0:
            q.put((rank, "init", rc, lib.sfc_last_error().decode()))
            return
        assert lib.sfc_comm_size(comm) == world and lib.sfc_comm_rank(comm) == rank
        ok = True
        for it in range(5):  # several generations: the two slot sets are reused
            mine = np.full(nbytes, (rank * 17 + it) % 251, dtype=np.uint8)
            mine[: 8] = np.frombuffer(np.int64(rank * 1000 + it).tobytes(), dtype=np.uint8)
            out = np.zeros(world * nbytes, dtype=np.uint8)
            assert lib.sfc_comm_allgather(comm, mine.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), nbytes) == 0
            for r in range(world):
                blk = out[r * nbytes:(r + 1) * nbytes]
                ok &= int(np.frombuffer(blk[:8].tobytes(), dtype=np.int64)[0]) == r * 1000 + it
                ok &= bool(np.all(blk[8:] == (r * 17 + it) % 251))
            if rank == it % world:
                time.sleep(0.02)  # a straggler must not let anybody run ahead by two generations
            assert lib.sfc_comm_barrier(comm) == 0
        # a host-only communicator has no device side: loud BackendError, no CPU fallback
        ptr = C.c_void_p()
        rc_alloc = lib.sfc_comm_alloc(comm, 1024, C.byref(ptr))
        dd = _lib.sfc_dist_desc()
        plan = C.c_void_p()
        rc_plan = lib.sfc_dist_plan_create(C.byref(plan), comm, C.byref(dd))
        lib.sfc_comm_destroy(comm)
        q.put((rank, "done", ok, (rc_alloc, rc_plan)))
    except Exception as ex:  # pragma: no cover
        q.put((rank, "exc", False, repr(ex)))


@pytest.mark.parametrize("world,nbytes", [(2, 64), (3, 1000), (4, 256)])
def test_rank_rendezvous_allgather_barrier(build_artifacts, world, nbytes):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    name = f"t{os.getpid()}_{world}_{nbytes}_{int(time.time() * 1e3) % 100000}"
    procs = [ctx.Process(target=_worker, args=(r, world, name, q, nbytes)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, what, ok, extra in res:
        assert what == "done" and ok, (rank, what, ok, extra)
        assert extra == (-6, -6), extra  # SFC_ERR_BACKEND twice
    assert not os.path.exists("/dev/shm/sfc_" + name)  # rank 0 unlinks the name once everybody is attached


def test_comm_argument_errors(build_artifacts, monkeypatch):
    from scirs_b200 import _lib

    lib = _lib.load()
    comm = C.c_void_p()
    assert lib.sfc_comm_init_rank(C.byref(comm), b"bad name!", 0, 2, -1) == _lib.SFC_ERR_VALUE
    assert lib.sfc_comm_init_rank(C.byref(comm), b"ok", 2, 2, -1) == _lib.SFC_ERR_VALUE
    assert lib.sfc_comm_init_rank(C.byref(comm), b"ok", 0, 17, -1) == _lib.SFC_ERR_VALUE
    # world size 1 needs no segment at all
    assert lib.sfc_comm_init_rank(C.byref(comm), b"solo", 0, 1, -1) == 0
    assert lib.sfc_comm_barrier(comm) == 0 and lib.sfc_comm_size(comm) == 1
    lib.sfc_comm_destroy(comm)
    # without a CUDA device the local (multi-GPU, one process) communicator fails loudly
    if lib.sfc_device_count() == 0:
        assert lib.sfc_comm_init_local(C.byref(comm), 2, None) == _lib.SFC_ERR_BACKEND
        assert lib.sfc_set_num_gpus(2) == _lib.SFC_ERR_VALUE
    assert lib.sfc_get_num_gpus() == 1


def test_rendezvous_times_out_when_a_rank_is_missing(build_artifacts):
    """A rank that never shows up must produce CommunicationError, not a hang."""
    import subprocess
    import sys

    code = (
        "import ctypes as C, sys\n"
        "sys.path.insert(0, %r)\n"
        "from scirs_b200 import _lib\n"
        "lib = _lib.load(); comm = C.c_void_p()\n"
        "rc = lib.sfc_comm_init_rank(C.byref(comm), b'lonely_%d', 0, 2, -1)\n"
        "print(rc, lib.sfc_last_error().decode())\n" % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.getpid())
    )
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, SFC_COMM_TIMEOUT_MS="300"), capture_output=True,
                       text=True, timeout=60)
    assert r.stdout.startswith("-8 "), r.stdout + r.stderr
    assert "did not reach the rendezvous" in r.stdout
    assert not os.path.exists("/dev/shm/sfc_lonely_%d" % os.getpid())
