"""Banded LDL' (cfg4: KKT matrices banded under the family's stage-interleaved ordering): kernels against NumPy,
the banded Newton-KKT step against the dense LDL' path, whole solves against the oracle."""

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from helpers import rel_err  # noqa: E402
from oracle import gradflow_oracle as orc  # noqa: E402
from pygradflow_b200 import synth  # noqa: E402


@pytest.fixture(scope="module", autouse=True)
def _gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


@pytest.mark.parametrize("N,bw", [(1, 1), (7, 3), (40, 23), (200, 23), (333, 63), (3072, 23)])
def test_band_factor_solve_vs_numpy(N, bw):
    from pygradflow_b200 import kernels as K
    from pygradflow_b200.kernels import WorkList

    B = 3
    rng = np.random.default_rng(N + bw)
    W = bw + 1
    dense = np.zeros((B, N, N))
    band = np.zeros((B, N, W))
    sign = np.where(rng.uniform(size=N) < 0.35, -1.0, 1.0)  # quasi-definite-like sign pattern, dominant diagonal
    for b in range(B):
        for t in range(N):
            for d in range(0, min(bw, t) + 1):
                v = sign[t] * (bw + 2.0 + rng.uniform()) if d == 0 else rng.uniform(-1, 1)
                dense[b, t, t - d] = dense[b, t - d, t] = v
                band[b, t, d] = v
    rhs = rng.standard_normal((B, N))
    f64 = dict(dtype=torch.float64, device="cuda")
    Kb, v = torch.as_tensor(band, **f64).contiguous(), torch.as_tensor(rhs, **f64).contiguous()
    info = torch.ones(B, dtype=torch.int32, device="cuda")
    nneg = torch.zeros(B, dtype=torch.int32, device="cuda")
    w = WorkList.all(B)
    K.band_factor(Kb, bw, info, nneg, w)
    K.band_solve(Kb, bw, v, w)
    fac = Kb.cpu().numpy()
    assert info.cpu().tolist() == [0] * B and nneg.cpu().tolist() == [int((sign < 0).sum())] * B
    for b in range(B):
        assert rel_err(v[b].cpu().numpy(), np.linalg.solve(dense[b], rhs[b])) <= 1e-12
        L = np.eye(N)
        for t in range(N):
            for d in range(1, min(bw, t) + 1):
                L[t, t - d] = fac[b, t, d]
        assert rel_err((L * fac[b, :, 0]) @ L.T, dense[b]) <= 1e-12


def test_band_factor_flags_zero_pivot():
    from pygradflow_b200 import kernels as K
    from pygradflow_b200.kernels import WorkList

    band = np.zeros((2, 6, 4))
    band[:, :, 0] = 2.0
    band[1, 3, 0] = 0.0
    Kb = torch.as_tensor(band, dtype=torch.float64, device="cuda")
    info = torch.zeros(2, dtype=torch.int32, device="cuda")
    nneg = torch.zeros(2, dtype=torch.int32, device="cuda")
    K.band_factor(Kb, 3, info, nneg, WorkList.all(2))
    assert info.cpu().tolist() == [0, 4]


def _ocp(B, S, nx, nu):
    from pygradflow_b200.problem import BatchedOCP

    d = synth.ocp_batch(range(B), stages=S, nx=nx, nu=nu)
    return BatchedOCP(d["A"], d["B"], d["Q"], d["R"], d["xinit"], d["umax"], d["h"]), d


@pytest.mark.parametrize("S,nx,nu", [(6, 3, 2), (16, 4, 3), (32, 8, 8)])
def test_banded_newton_step_matches_dense(S, nx, nu):
    """One Newton-KKT step (active controls included) through the banded engine and through the dense LDL' engine."""
    from pygradflow_b200.newton import NewtonKKTStepper
    from pygradflow_b200.params import LinearSolverType

    B = 4
    prob, d = _ocp(B, S, nx, nu)
    order, bw = prob.kkt_band()
    assert sorted(order) == list(range(prob.n + prob.m)) and bw == nx + nx + nu - 1
    rng = np.random.default_rng(1)
    f64 = dict(dtype=torch.float64, device="cuda")
    x = torch.as_tensor(np.clip(0.6 * rng.standard_normal((B, prob.n)), -0.4, 0.4), **f64)  # many controls at a bound
    y = torch.as_tensor(0.3 * rng.standard_normal((B, prob.m)), **f64)
    lamb = torch.as_tensor(10.0 ** rng.uniform(-1, 1, B), **f64)
    rho = torch.as_tensor(10.0 ** rng.uniform(-4, 0, B), **f64)
    out = {}
    for lin in (LinearSolverType.Banded, LinearSolverType.LDLT):
        st = NewtonKKTStepper(prob, lin)
        assert st.engine.linear == lin
        xn, yn, diff, fn, info = st.step(x, y, lamb, rho)
        out[lin] = (xn.clone(), yn.clone(), diff.clone(), info.clone(), st.engine.active.clone())
    a, b_ = out[LinearSolverType.Banded], out[LinearSolverType.LDLT]
    assert int(a[3].abs().sum().item()) == 0 and int(b_[3].abs().sum().item()) == 0
    assert torch.equal(a[4], b_[4])
    assert rel_err(a[0].cpu().numpy(), b_[0].cpu().numpy()) <= 1e-11
    assert rel_err(a[1].cpu().numpy(), b_[1].cpu().numpy()) <= 1e-11
    assert rel_err(a[2].cpu().numpy(), b_[2].cpu().numpy()) <= 1e-11


@pytest.mark.parametrize("S,nx,nu,B", [(6, 3, 2, 5), (16, 4, 3, 4)])
def test_banded_ocp_solve_vs_oracle(S, nx, nu, B):
    from pygradflow_b200.params import LinearSolverType, Params
    from pygradflow_b200.solver import BatchedSolver

    prob, d = _ocp(B, S, nx, nu)
    solver = BatchedSolver(prob, Params(linear_solver_type=LinearSolverType.Banded))
    assert solver.engine.linear == LinearSolverType.Banded and solver.engine.K is None
    res = solver.solve(d["x0"], d["y0"])
    for b in range(B):
        p = orc.OCP(d["A"][b], d["B"][b], d["Q"][b], d["R"][b], d["xinit"][b], d["umax"], d["h"])
        ref = orc.Solver(p, orc.OracleParams()).solve(d["x0"][b], d["y0"][b])
        assert int(res.status[b].item()) == ref.status == 1
        assert rel_err(res.x[b].cpu().numpy(), ref.x) <= 1e-5
        assert abs(int(res.iterations[b].item()) - ref.iterations) <= 2
