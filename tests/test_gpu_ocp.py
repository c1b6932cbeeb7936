"""cfg4 family (discretised nonlinear optimal control: equality dynamics + control bounds) on the B200 path:
device evaluators against the oracle's OCP class, batched solves against the oracle and against traces of the
REAL reference (tests/golden/ocp.npz)."""

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from helpers import noise_horizon, rel_err  # noqa: E402
from oracle import gradflow_oracle as orc  # noqa: E402
from pygradflow_b200 import synth  # noqa: E402


@pytest.fixture(scope="module", autouse=True)
def _gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def _batch(B, S, nx, nu):
    from pygradflow_b200.problem import BatchedOCP

    d = synth.ocp_batch(range(B), stages=S, nx=nx, nu=nu)
    prob = BatchedOCP(d["A"], d["B"], d["Q"], d["R"], d["xinit"], d["umax"], d["h"])
    refs = [orc.OCP(d["A"][b], d["B"][b], d["Q"][b], d["R"][b], d["xinit"][b], d["umax"], d["h"]) for b in range(B)]
    return prob, refs, d


@pytest.mark.parametrize("S,nx,nu", [(6, 3, 2), (16, 4, 3), (32, 8, 8)])
def test_ocp_evaluators_vs_oracle(S, nx, nu):
    from pygradflow_b200.kernels import WorkList

    B = 5
    prob, refs, d = _batch(B, S, nx, nu)
    n, m = prob.n, prob.m
    assert (n, m) == (S * (nx + nu), S * nx)
    rng = np.random.default_rng(11)
    z = 0.5 * rng.standard_normal((B, n))
    y = rng.standard_normal((B, m))
    f64 = dict(dtype=torch.float64, device="cuda")
    zt, yt = torch.as_tensor(z, **f64), torch.as_tensor(y, **f64)
    grad, cons, obj = torch.zeros((B, n), **f64), torch.zeros((B, m), **f64), torch.zeros((B,), **f64)
    J, H = torch.full((B, m, n), 7.0, **f64), torch.full((B, n, n), 7.0, **f64)  # must be zeroed by the family
    w = WorkList.all(B)
    prob.eval(zt, grad, cons, obj, w)
    Jo = prob.jac(zt, J, w)
    Ho = prob.lag_hess(zt, yt, H, w)
    for b in range(B):
        p = refs[b]
        assert np.array_equal(grad[b].cpu().numpy(), p.obj_grad(z[b]))          # products only: bit-exact
        assert rel_err(cons[b].cpu().numpy(), p.cons(z[b])) <= 1e-14            # sin(): last-bit differences
        assert abs(obj[b].item() - p.obj(z[b])) <= 1e-12 * max(1.0, abs(p.obj(z[b])))
        assert rel_err(Jo[b].cpu().numpy(), p.cons_jac(z[b])) <= 1e-14
        assert rel_err(Ho[b].cpu().numpy(), p.lag_hess(z[b], y[b])) <= 1e-14
    assert np.array_equal(prob.var_lb[0].cpu().numpy(), refs[0].var_lb)
    assert np.array_equal(prob.var_ub[0].cpu().numpy(), refs[0].var_ub)


def _trace_solve(prob, params, x0, y0):
    from pygradflow_b200.solver import BatchedSolver

    solver = BatchedSolver(prob, params)
    B = prob.B
    traces = [[] for _ in range(B)]

    def hook(outer, s):
        ph, st = s.phase.cpu().numpy(), s.status.cpu().numpy()
        pts = {2: s.mid, 3: s.fin, 4: s.fin}
        th = s.theta.cpu().numpy()
        for b in range(B):
            if st[b] != 0 or ph[b] == 0:
                continue
            src = pts.get(int(ph[b]), s.cur)
            traces[b].append(dict(x=src[0][b].cpu().numpy().copy(), y=src[1][b].cpu().numpy().copy(),
                                  accept=ph[b] in (2, 3), theta=float(th[b]) if ph[b] in (3, 4) else float("nan")))

    return solver.solve(x0, y0, on_iteration=hook), traces


@pytest.mark.parametrize("newton", ["Simplified", "Full"])
@pytest.mark.parametrize("S,nx,nu,B", [(6, 3, 2, 6), (16, 4, 3, 4)])
def test_batched_ocp_vs_oracle(S, nx, nu, B, newton):
    from pygradflow_b200.params import NewtonType, Params

    prob, refs, d = _batch(B, S, nx, nu)
    res, traces = _trace_solve(prob, Params(newton_type=NewtonType[newton]), d["x0"], d["y0"])
    for b in range(B):
        ref = orc.Solver(refs[b], orc.OracleParams(newton_type=newton.lower())).solve(d["x0"][b], d["y0"][b], record=True)
        assert int(res.status[b].item()) == ref.status == 1
        horizon = min(noise_horizon(ref.trace), noise_horizon(traces[b]))
        got, exp = [t["accept"] for t in traces[b]], [t["accept"] for t in ref.trace]
        assert got[: horizon + 1] == exp[: horizon + 1]
        for i in range(min(horizon + 1, len(exp), len(got))):
            tol = 1e-10 if i < 8 else 1e-7  # long nonlinear trajectories amplify last-bit differences
            assert rel_err(traces[b][i]["x"], ref.trace[i]["x"]) <= tol, (b, i)
            assert rel_err(traces[b][i]["y"], ref.trace[i]["y"]) <= tol, (b, i)
        if horizon >= len(exp):
            assert int(res.iterations[b].item()) == ref.iterations
            assert int(res.accepted_steps[b].item()) == ref.accepted_steps
        assert rel_err(res.x[b].cpu().numpy(), ref.x) <= 1e-5


@pytest.mark.parametrize("newton", ["Simplified", "Full"])
def test_batched_ocp_golden_reference(golden, newton):
    """Traces of the real reference Solver.solve on the cfg4 family."""
    from pygradflow_b200.params import NewtonType, Params
    from pygradflow_b200.problem import BatchedOCP

    g = golden("ocp")
    for (S, nx, nu, k) in [(6, 3, 2, 0), (16, 4, 3, 1)]:
        key = f"ocp_S{S}_nx{nx}_nu{nu}_k{k}/{newton}"
        d = synth.ocp_batch([k], stages=S, nx=nx, nu=nu)
        prob = BatchedOCP(d["A"], d["B"], d["Q"], d["R"], d["xinit"], d["umax"], d["h"])
        res, traces = _trace_solve(prob, Params(newton_type=NewtonType[newton]), d["x0"], d["y0"])
        assert int(res.status[0].item()) == int(g[f"{key}/status"]) == 1
        horizon = noise_horizon(traces[0])
        accepts = list(g[f"{key}/accepts"])
        assert [t["accept"] for t in traces[0]][: horizon + 1] == accepts[: horizon + 1]
        for row, i in enumerate(g[f"{key}/trace_idx"]):
            if i <= horizon and i < len(traces[0]):
                tol = 1e-10 if i < 8 else 1e-7
                assert rel_err(traces[0][i]["x"], g[f"{key}/trace_x"][row]) <= tol, (key, i)
        if horizon >= len(accepts):
            assert int(res.iterations[0].item()) == int(g[f"{key}/iterations"])
            assert int(res.accepted_steps[0].item()) == int(g[f"{key}/accepted_steps"])
        assert rel_err(res.x[0].cpu().numpy(), g[f"{key}/x"]) <= 1e-5


# ------------------------------------------------------------------ stage-structured engine (gf_blocktri.cu)
@pytest.mark.parametrize("S,nu", [(1, 8), (2, 4), (5, 8), (16, 4), (37, 16), (128, 8)])
def test_stage_layout_and_step_vs_oracle(S, nu):
    """Compact Jacobian / diagonal Hessian equal the dense ones entry by entry; J'v through the compact layout; one
    Newton-KKT step of the stage-structured engine (Schur complement on the multipliers + block cyclic reduction)
    against the oracle's dense SymmetricStepSolver: x, y within 1e-10, identical active sets."""
    from pygradflow_b200 import kernels as K
    from pygradflow_b200.kernels import WorkList
    from pygradflow_b200.newton import NewtonKKTStepper
    from pygradflow_b200.params import LinearSolverType

    B, nx = 4, 8
    prob, refs, d = _batch(B, S, nx, nu)
    n, m = prob.n, prob.m
    rng = np.random.default_rng(5 + S)
    z = 0.6 * rng.standard_normal((B, n))
    z[:, np.arange(n) % (nx + nu) >= nx] = np.clip(z[:, np.arange(n) % (nx + nu) >= nx], -d["umax"], d["umax"])
    y = 0.5 * rng.standard_normal((B, m))
    f64 = dict(dtype=torch.float64, device="cuda")
    zt, yt = torch.as_tensor(z, **f64), torch.as_tensor(y, **f64)
    w = WorkList.all(B)
    Jd = prob.jac(zt, torch.zeros((B, m, n), **f64), w).clone()
    Hd = prob.lag_hess(zt, yt, torch.zeros((B, n, n), **f64), w).clone()
    st = NewtonKKTStepper(prob, LinearSolverType.Auto)
    assert st.engine.linear == LinearSolverType.BlockTri and prob.compact
    Jc = prob.jac(zt, prob.alloc_jac(), w)
    Hc = prob.lag_hess(zt, yt, prob.alloc_hess(), w)
    wv = nx + nu
    assert Jc.shape == (B, S, nx + wv, nx)
    for j in range(S):   # Jc[b][j][c][r] = d c_{j,r} / d (column c): transposed stage blocks of the dense Jacobian
        rows = slice(j * nx, (j + 1) * nx)
        assert torch.equal(Jc[:, j, nx:, :].transpose(1, 2), Jd[:, rows, j * wv:(j + 1) * wv])
        if j >= 1:
            assert torch.equal(Jc[:, j, :nx, :].transpose(1, 2), Jd[:, rows, (j - 1) * wv:(j - 1) * wv + nx])
    z2 = zt + 0.1   # a second evaluation into the same buffer rewrites only the x-dependent entries
    Jc2 = prob.jac(z2, Jc, w).clone()
    prob._zeroed.clear()
    assert torch.equal(Jc2, prob.jac(z2, prob.alloc_jac(), w))
    prob.jac(zt, Jc, w)
    assert torch.equal(Hc, torch.diagonal(Hd, dim1=1, dim2=2))
    grad, cons = torch.as_tensor(rng.standard_normal((B, n)), **f64), torch.as_tensor(rng.standard_normal((B, m)), **f64)
    rho = torch.as_tensor(10.0 ** rng.uniform(-6, 0, B), **f64)
    outs = []
    for J, fn in ((Jd, K.aug_lag_grad), (Jc, prob.aug_lag_grad)):
        dL, jty, jtc = (torch.zeros((B, n), **f64) for _ in range(3))
        fn(J, grad, cons, yt, rho, dL, jty, jtc, w)
        outs.append((dL, jty, jtc))
    for a, b_ in zip(*outs):
        assert rel_err(b_.cpu().numpy(), a.cpu().numpy()) <= 1e-13
    lamb = torch.as_tensor(10.0 ** rng.uniform(-2, 1.5, B), **f64)
    xn, yn, diff, fn_, info = st.step(zt, yt, lamb, rho)
    torch.cuda.synchronize()
    assert int((info != 0).sum().item()) == 0
    for b in range(B):
        prm = orc.OracleParams(newton_type="full")
        it = orc.Iterate(refs[b], prm, z[b], y[b])
        res = orc.newton_method(refs[b], prm, it, 1.0 / float(lamb[b].item()), float(rho[b].item())).step(it)
        assert np.array_equal(st.engine.active[b].cpu().numpy().astype(bool), res.active_set)
        assert rel_err(xn[b].cpu().numpy(), res.iterate.x) <= 1e-10, (S, b)
        assert rel_err(yn[b].cpu().numpy(), res.iterate.y) <= 1e-10, (S, b)
        assert abs(diff[b].item() - res.diff) <= 1e-10 * max(1.0, res.diff)


def test_stage_engine_flags_indefinite_hessian():
    """H_ii + lamb <= 0 for an inactive variable: K is not quasi-definite, info = -2 (the step is rejected)."""
    from pygradflow_b200.kernels import WorkList
    from pygradflow_b200.newton import NewtonKKTStepper
    from pygradflow_b200.params import LinearSolverType

    B, S, nx, nu = 3, 8, 8, 8
    prob, refs, d = _batch(B, S, nx, nu)
    prob.Q[1, 2, 3] = -5.0
    st = NewtonKKTStepper(prob, LinearSolverType.BlockTri)
    f64 = dict(dtype=torch.float64, device="cuda")
    z, y = torch.zeros((B, prob.n), **f64), torch.zeros((B, prob.m), **f64)
    out = st.step(z, y, torch.full((B,), 1.0, **f64), torch.full((B,), 1e-3, **f64))
    assert out[4].cpu().tolist() == [0, -2, 0]
