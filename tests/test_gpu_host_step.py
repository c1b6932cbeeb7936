"""Host-buffer entry (what bench.py times as e2e): chunked, double-buffered Newton-KKT steps with pinned host inputs
give bit-identical results to the device-resident stepper; the symmetric Hessian crosses PCIe as its lower block
triangle only and is rebuilt exactly on the device."""

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from pygradflow_b200 import synth  # noqa: E402


@pytest.fixture(scope="module", autouse=True)
def _gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


@pytest.mark.parametrize("n", [32, 64, 100, 130, 512])
def test_symmetric_transfer_rebuilds_hessian(n):
    from pygradflow_b200 import kernels as K

    B = 5
    rng = np.random.default_rng(n)
    M = rng.standard_normal((B, n, n))
    H = M + M.transpose(0, 2, 1)
    host = torch.as_tensor(H).pin_memory()
    dev = torch.full((B + 1, n, n), float("nan"), dtype=torch.float64, device="cuda")
    K.h2d_sym_lower(dev, host, B)
    K.symmetrize_lower(dev, B)
    torch.cuda.synchronize()
    assert np.array_equal(dev[:B].cpu().numpy(), H)
    assert torch.isnan(dev[B]).all()  # nothing written past cnt
    assert K.h2d_sym_lower_bytes(B, n) <= 8 * B * n * n


@pytest.mark.parametrize("n,m,B,chunk", [(130, 40, 7, 3), (64, 0, 5, 2)])
def test_host_step_matches_device_stepper(n, m, B, chunk):
    from pygradflow_b200.host_step import HostNewtonKKT
    from pygradflow_b200.newton import NewtonKKTStepper
    from pygradflow_b200.problem import BatchedQP

    d = synth.qp_batch(range(B), n, m)
    rng = np.random.default_rng(5)
    x = np.clip(0.3 * rng.uniform(-1, 1, (B, n)), -1, 1)
    y = 0.1 * rng.standard_normal((B, m))
    lamb, rho = 10.0 ** rng.uniform(-1, 1, B), 10.0 ** rng.uniform(-4, 0, B)
    f64 = dict(dtype=torch.float64, device="cuda")
    prob = BatchedQP(d["H"], d["A"] if m else None, d["g"], d["b"] if m else None, d["lb"], d["ub"])
    st = NewtonKKTStepper(prob)
    xn, yn, diff, fn, info = (t.clone() for t in st.step(torch.as_tensor(x, **f64), torch.as_tensor(y, **f64),
                                                         torch.as_tensor(lamb, **f64), torch.as_tensor(rho, **f64)))
    hk = HostNewtonKKT(n, m, chunk=chunk)
    pin = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64).pin_memory()
    host = dict(H=pin(d["H"]), A=pin(d["A"]), g=pin(d["g"]), b=pin(d["b"]), lb=pin(d["lb"]), ub=pin(d["ub"]), x=pin(x),
                y=pin(y), lamb=pin(lamb), rho=pin(rho))
    out = dict(xn=torch.empty((B, n), dtype=torch.float64).pin_memory(), yn=torch.empty((B, m), dtype=torch.float64).pin_memory(),
               diff=torch.empty((B,), dtype=torch.float64).pin_memory(), fnorm=torch.empty((B,), dtype=torch.float64).pin_memory(),
               info=torch.empty((B,), dtype=torch.int32).pin_memory())
    from pygradflow_b200.host_step import ResidentNewtonKKT

    rk = ResidentNewtonKKT.register({k: host[k] for k in ("H", "A", "g", "b", "lb", "ub") if m or k not in ("A", "b")})
    rk.step(host, out)  # problem registered once; only x, y, lamb, rho cross the bus
    torch.cuda.synchronize()
    assert np.array_equal(out["xn"].numpy(), xn.cpu().numpy()) and np.array_equal(out["diff"].numpy(), diff.cpu().numpy())
    assert np.array_equal(out["fnorm"].numpy(), fn.cpu().numpy()) and np.array_equal(out["info"].numpy(), info.cpu().numpy())
    assert rk.bytes_per_step()[0] == 8 * B * (n + m + 2)
    for t in out.values():
        t.zero_()
    for _ in range(2):  # the second pass reuses the slots
        hk.step(host, out)
        torch.cuda.synchronize()
        assert np.array_equal(out["xn"].numpy(), xn.cpu().numpy())
        if m:
            assert np.array_equal(out["yn"].numpy(), yn.cpu().numpy())
        assert np.array_equal(out["diff"].numpy(), diff.cpu().numpy())
        assert np.array_equal(out["fnorm"].numpy(), fn.cpu().numpy())
        assert np.array_equal(out["info"].numpy(), info.cpu().numpy())
    h2d, d2h = hk.bytes_per_step(B)
    assert h2d <= 8 * B * (n * n + m * n + 4 * n + 2 * m + 2) and d2h == 8 * B * (n + m + 2) + 4 * B
