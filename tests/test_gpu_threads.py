"""GPU: concurrent use of the library from several host threads and several CUDA streams.

The reference's free functions are safe to call from rayon workers (fft/planning.rs:124-138 shares one
`Arc<dyn Fft>` across threads); equal descriptors resolve to ONE cached plan here, whose multi-pass flavours
(four-step, Bluestein, Hermitian fill) own device scratch — so two threads, or two streams, must not run over
the same scratch at the same time.  Plan::exec orders an execution on another stream after the previous one."""
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = 1e-12


@pytest.fixture(scope="module")
def sb(build_artifacts):
    import scirs_b200 as m
    from scirs_b200 import _lib

    lib = _lib.load()
    assert lib.sfc_device_count() >= 1
    m.error.check(lib.sfc_init(0))
    return m


def _rel(a, b):
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


@pytest.mark.parametrize("n", [1 << 20, 100003, 3 ** 9])
def test_two_threads_same_shape_free_functions(sb, n):
    """Each thread has its own thread_local workspace stream but both resolve to the same cached plan."""
    from scirs_b200 import _lib

    lib = _lib.load()
    rng = np.random.default_rng(n)
    xs = [rng.standard_normal(n) + 1j * rng.standard_normal(n) for _ in range(4)]
    refs = [np.fft.fft(x) for x in xs]
    errs, fails = [], []

    def work(tid):
        try:
            sb.error.check(lib.sfc_init(0))
            for it in range(6):
                k = (tid * 2 + it) % 4
                got = sb.fft(xs[k], n)
                errs.append(_rel(got, refs[k]))
        except Exception as ex:  # pragma: no cover
            fails.append(repr(ex))

    ts = [threading.Thread(target=work, args=(t,)) for t in range(2)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not fails, fails
    assert len(errs) == 12 and max(errs) < TOL, max(errs)


@pytest.mark.parametrize("n", [1 << 20, 100003])
def test_two_handles_two_streams(sb, n):
    """Two sfc_plan handles with equal descriptors (one shared cached plan) executed back to back on two streams."""
    import torch

    dev = torch.device("cuda:0")
    b = 4
    g = torch.Generator(device=dev).manual_seed(n)
    x = [torch.view_as_complex(torch.randn(b, n, 2, dtype=torch.float64, device=dev, generator=g)) for _ in range(2)]
    y = [torch.empty_like(x[0]) for _ in range(2)]
    plans = [sb.FftPlan([b, n], [1]) for _ in range(2)]
    streams = [torch.cuda.Stream(dev) for _ in range(2)]
    torch.cuda.synchronize()
    for rep in range(5):
        for i in range(2):
            plans[i].execute_device(x[i], y[i], streams[i].cuda_stream)
    torch.cuda.synchronize()
    for i in range(2):
        ref = np.fft.fft(x[i].cpu().numpy(), axis=1)
        assert _rel(y[i].cpu().numpy(), ref) < TOL


def test_scratch_is_lazy_and_released_under_pressure(sb):
    """Cached plans hold no scratch until they run, and idle plans give theirs back when an allocation fails
    (up to 128 cached plans x ~2 GiB each could exhaust HBM)."""
    p = sb.FftPlan([8, 1 << 21], [1])
    assert p.info["scratch_bytes"] > 0
    x = np.zeros((8, 1 << 21), dtype=np.complex128)
    x[:, 1] = 1.0
    got = p.execute(x).reshape(8, -1)
    k = np.arange(1 << 21)
    assert _rel(got[3], np.exp(-2j * np.pi * k / (1 << 21))) < TOL
