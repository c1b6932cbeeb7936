"""Batched slack transform (cons_problem.py / transform.py) in front of the B200 Newton/KKT path:
wrapper callbacks against the oracle's ConstrainedProblem, whole solves against the oracle and the REAL reference."""

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


def _batch(B, n, m):
    from pygradflow_b200.problem import BatchedQP

    d = synth.general_qp_batch(range(B), n, m)
    prob = BatchedQP(d["H"], d["A"], d["g"], d["b"], d["lb"], d["ub"])
    refs = [orc.GeneralQP(d["H"][b], d["A"][b], d["g"][b], d["b"][b], d["lb"][b], d["ub"][b], d["cons_lb"][b],
                          d["cons_ub"][b]) for b in range(B)]
    return prob, refs, d


def test_constrained_wrapper_callbacks_vs_oracle():
    from pygradflow_b200.kernels import WorkList
    from pygradflow_b200.transform import BatchedConstrained

    B, n, m = 4, 16, 8
    prob, refs, d = _batch(B, n, m)
    cp = BatchedConstrained(prob, d["cons_lb"], d["cons_ub"])
    ocp = [orc.ConstrainedProblem(r) for r in refs]
    ns = cp.ns
    assert ns == len(ocp[0].slack_positions) == m - m // 2 and (cp.n, cp.m) == (n + ns, m)
    rng = np.random.default_rng(2)
    x, y = rng.standard_normal((B, n + ns)), rng.standard_normal((B, m))
    f64 = dict(dtype=torch.float64, device="cuda")
    xt, yt = torch.as_tensor(x, **f64), torch.as_tensor(y, **f64)
    grad, cons, obj = torch.zeros((B, n + ns), **f64), torch.zeros((B, m), **f64), torch.zeros((B,), **f64)
    w = WorkList.all(B)
    cp.eval(xt, grad, cons, obj, w)
    J = cp.jac(xt, torch.zeros((B, m, n + ns), **f64), w)
    H = cp.lag_hess(xt, yt, torch.zeros((B, n + ns, n + ns), **f64), w)
    for b in range(B):
        assert rel_err(grad[b].cpu().numpy(), ocp[b].obj_grad(x[b])) <= 1e-13
        assert rel_err(cons[b].cpu().numpy(), ocp[b].cons(x[b])) <= 1e-13
        assert np.array_equal(J[b].cpu().numpy(), ocp[b].cons_jac(x[b]))
        assert np.array_equal(H[b].cpu().numpy(), ocp[b].lag_hess(x[b], y[b]))
        assert np.array_equal(cp.var_lb[b].cpu().numpy(), ocp[b].var_lb)
        assert np.array_equal(cp.var_ub[b].cpu().numpy(), ocp[b].var_ub)
    x0 = torch.as_tensor(d["x0"], **f64)
    xs, _ = cp.transform_sol(x0, torch.zeros((B, m), **f64))
    for b in range(B):
        assert rel_err(xs[b].cpu().numpy(), ocp[b].transform_sol(d["x0"][b], np.zeros(m))[0]) <= 1e-13


@pytest.mark.parametrize("n,m,B", [(16, 8, 6), (24, 12, 4)])
def test_solve_general_vs_oracle(n, m, B):
    from pygradflow_b200.transform import solve_general

    prob, refs, d = _batch(B, n, m)
    res = solve_general(prob, d["cons_lb"], d["cons_ub"], None, d["x0"], d["y0"])
    assert res.x.shape == (B, n)
    for b in range(B):
        ref = orc.solve_general(refs[b], orc.OracleParams(), d["x0"][b], d["y0"][b])
        assert int(res.status[b].item()) == ref.status == 1
        assert rel_err(res.x[b].cpu().numpy(), ref.x) <= 1e-4  # both stop at opt_tol = 1e-6 on their own noise path
        c = refs[b].cons(res.x[b].cpu().numpy())
        assert (c >= d["cons_lb"][b] - 1e-6).all() and (c <= d["cons_ub"][b] + 1e-6).all()
        assert abs(int(res.iterations[b].item()) - ref.iterations) <= 4  # past the rounding-noise horizon lambda differs (the summation order of the LU substitution decides)


def test_solve_general_golden_reference(golden):
    """The reference's Solver.solve on a general QP (it applies ConstrainedProblem itself)."""
    from pygradflow_b200.transform import solve_general

    g = golden("constrained")
    prob, refs, d = _batch(1, 16, 8)
    res = solve_general(prob, d["cons_lb"], d["cons_ub"], None, d["x0"], d["y0"])
    key = "gqp_n16_m8_k0/Simplified"
    assert int(res.status[0].item()) == int(g[f"{key}/status"])
    assert rel_err(res.x[0].cpu().numpy(), g[f"{key}/x"]) <= 1e-4
    assert abs(int(res.iterations[0].item()) - int(g[f"{key}/iterations"])) <= 2
