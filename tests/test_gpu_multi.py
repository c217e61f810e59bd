"""GPU: the multi-GPU paths behind the C ABI (csrc/dist.cu, SURVEY 8e / 8b last row) on real hardware.

On a box with >= 2 GPUs this spawns one process per GPU (no torchrun, no torch.distributed: the library does its own
rendezvous) on 2 GPUs and on all visible GPUs and checks the slab-decomposed fftn — both layouts, both directions,
f64 and f32, three consecutive calls (both receive buffers) — against the oracle / the single-GPU plan with the
north_star tolerance (rel-L2 <= 1e-12 f64, 1e-5 f32); the batch split; the one-process-many-GPUs mode; the free
functions under sfc_set_num_gpus; and the Python-free C++ driver.  With one GPU the P = 1 degenerate forms run.

Reference: the decomposition tests of scirs2-fft/src/distributed.rs:844-1001 check shapes only — its exchange is a mock."""
import ctypes as C
import json
import os
import subprocess
import sys
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    from scirs_b200 import _lib

    return _lib.load().sfc_device_count()


def _worlds():
    n = _ngpu()
    w = [p for p in (2, 4, 8) if p <= n]
    if os.environ.get("SFC_TEST_WORLDS"):  # e.g. "8": only that world size (GPU-minutes on a big box are expensive)
        w = [int(v) for v in os.environ["SFC_TEST_WORLDS"].split(",") if int(v) <= n]
    return w or [1]


def _spawn(world, case, sizes, timeout=900):
    name = f"pt{os.getpid()}_{case}_{world}_{int(time.time() * 1e3) % 1000000}"
    procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "dist_worker.py"), "--rank", str(r), "--world", str(world),
                               "--name", name, "--case", case, "--sizes", sizes], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                              text=True, env=dict(os.environ, SFC_COMM_TIMEOUT_MS="120000")) for r in range(world)]
    outs = []
    for p in procs:
        try:
            o, e = p.communicate(timeout=timeout)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        outs.append((p.returncode, o, e))
    records = []
    for rc, o, e in outs:
        assert rc == 0, o[-3000:] + e[-3000:]
        records += [json.loads(l) for l in o.splitlines() if l.startswith("{")]
    assert records and all(r["ok"] for r in records)
    return records


@pytest.mark.parametrize("world", _worlds())
def test_slab_and_batch_split_one_process_per_gpu_host_buffers(build_artifacts, world):
    recs = _spawn(world, "host", "64,256")
    worst64 = max(r["rel_l2"] for r in recs if r["prec"] == "f64")
    worst32 = max(r["rel_l2"] for r in recs if r["prec"] == "f32")
    print(f"world {world}: {len(recs)} checks, worst rel-L2 f64 {worst64:.2e} f32 {worst32:.2e}")
    assert worst64 <= 1e-12 and worst32 <= 1e-5
    assert {r["layout"] for r in recs} >= {"transposed", "natural", "batch_split"}


@pytest.mark.parametrize("world", _worlds())
def test_slab_one_process_per_gpu_device_buffers_and_symmetric_output(build_artifacts, world):
    recs = _spawn(world, "device", "128,512")
    assert max(r["rel_l2"] for r in recs) <= 1e-12
    if world > 1:
        assert any(r["layout"] == "natural:symmetric" for r in recs)


def test_one_process_many_gpus_and_free_functions(build_artifacts):
    """Local mode: sfc_comm_init_local, whole-array host execution, per-GPU device pointers, sfc_set_num_gpus."""
    import torch

    import scirs_b200 as sb
    from oracle import scirs2_fft_oracle as orc
    from scirs_b200.distributed import Communicator, DistPlan, get_num_gpus, set_num_gpus

    P = max(w for w in _worlds())
    sb.error.check(sb._lib.load().sfc_init(0))
    rng = np.random.default_rng(21)
    comm = Communicator.local(P)
    assert comm.size() == P
    shape = (64, 128, 32)
    x = rng.standard_normal(shape) + 1j * rng.standard_normal(shape)
    ref = orc.fftn(x)
    plan = DistPlan(comm, shape, [0, 1, 2], "slab", "natural")
    out = np.zeros(shape, dtype=np.complex128)
    for _ in range(2):
        out[...] = 0
        plan.execute_host(x, out)
        assert orc.rel_l2(out, ref) <= 1e-12
    # device pointers, one per GPU; transposed and natural layouts
    s0, s1 = shape[0] // P, shape[1] // P
    for layout in ("transposed", "natural"):
        pl = DistPlan(comm, shape, [0, 1, 2], "slab", layout)
        ins = [torch.from_numpy(np.ascontiguousarray(x[r * s0:(r + 1) * s0])).to(f"cuda:{r}") for r in range(P)]
        outs = [torch.zeros(pl.local_out_shape, dtype=torch.complex128, device=f"cuda:{r}") for r in range(P)]
        for r in range(P):
            torch.cuda.synchronize(r)
        for _ in range(3):
            pl.execute_device_multi(ins, outs)
        pl.synchronize()
        for r in range(P):
            want = ref[:, r * s1:(r + 1) * s1, :] if layout == "transposed" and P > 1 else ref[r * s0:(r + 1) * s0]
            assert orc.rel_l2(outs[r].cpu().numpy(), want) <= 1e-12, (layout, r)
        pl.close()
    # batch split over the GPUs of this process, ragged
    B, n = 5 * P + 3, 4096
    sig = rng.standard_normal((B, n)) + 1j * rng.standard_normal((B, n))
    bp = DistPlan(comm, [B, n], [1], "batch_split")
    got = np.zeros((B, n), dtype=np.complex128)
    bp.execute_host(sig, got)
    assert orc.rel_l2(got, np.fft.fft(sig, axis=1)) <= 1e-12
    bp.close()
    plan.close()
    comm.close()
    # the drop-in free functions over several GPUs: no new arguments, same results
    try:
        set_num_gpus(P)
        assert get_num_gpus() == P
        big = rng.standard_normal((128, 64, 64)) + 1j * rng.standard_normal((128, 64, 64))
        assert orc.rel_l2(sb.fftn(big), orc.fftn(big)) <= 1e-12
        assert orc.rel_l2(sb.ifftn(big, None, [2, 0, 1], "ortho"), orc.ifftn(big, None, [2, 0, 1], "ortho")) <= 1e-12
        odd = rng.standard_normal((6, 10, 12)) + 0j  # not a slab shape: runs on one GPU, same answer
        assert orc.rel_l2(sb.fftn(odd), orc.fftn(odd)) <= 1e-12
        ex = sb.planning.ParallelExecutor([1000], True)
        rows = [rng.standard_normal(1000) + 1j * rng.standard_normal(1000) for _ in range(4 * P + 1)]
        outs = [np.zeros(1000, dtype=np.complex128) for _ in rows]
        ex.execute_batch(rows, outs)
        for a, b in zip(rows, outs):
            assert orc.rel_l2(b, np.fft.fft(a)) <= 1e-12
    finally:
        set_num_gpus(1)


def test_cpp_driver_without_python_in_the_data_path(build_artifacts):
    exe = os.path.join(ROOT, "build", "cpp_dist_test")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    libdir = os.path.join(ROOT, "scirs_b200", "lib")
    subprocess.run(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp_dist_test.cpp"),
                    "-o", exe, "-L", libdir, "-lscirs2_fft_cuda", f"-Wl,-rpath,{libdir}"], check=True)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and "cpp dist ok" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]


def test_exchange_wait_is_bounded(build_artifacts):
    """A peer that never signals must end in CommunicationError after the timeout, not in a hung GPU."""
    if _ngpu() < 2:
        pytest.skip("needs two GPUs")
    code = r'''
import ctypes as C, os, sys, time
sys.path.insert(0, %r)
import numpy as np
from scirs_b200 import _lib
from scirs_b200.distributed import Communicator, DistPlan
rank = int(sys.argv[1])
lib = _lib.load(); lib.sfc_init(rank)
comm = Communicator.rank_mode(sys.argv[2], rank, 2, rank)
plan = DistPlan(comm, (64, 64, 64), [0, 1, 2], "slab", "transposed")
x = np.ones((32, 64, 64), dtype=np.complex128); out = np.zeros((64, 32, 64), dtype=np.complex128)
if rank == 1:
    time.sleep(6)   # never executes: rank 0 must give up on its own
    os._exit(0)
try:
    plan.execute_host(x, out)
    print("NO ERROR")
except Exception as ex:
    print("ERR", type(ex).__name__, ex)
os._exit(0)
''' % ROOT
    name = f"to{os.getpid()}_{int(time.time())}"
    env = dict(os.environ, SFC_EXCHANGE_TIMEOUT_MS="1500")
    ps = [subprocess.Popen([sys.executable, "-c", code, str(r), name], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env)
          for r in range(2)]
    o0, e0 = ps[0].communicate(timeout=120)
    ps[1].communicate(timeout=120)
    assert "ERR CommunicationError" in o0, o0 + e0
