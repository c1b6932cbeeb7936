"""Batched power-of-two scaling (scale.py) in front of the B200 Newton/KKT path: the weights of
create_scaling and the scaled callbacks bit for bit against the oracle's restatement (which is pinned to the REAL
reference by tests/golden/scaling.npz), whole solves against the oracle and the reference's results."""

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from helpers import rel_err  # noqa: E402
from oracle import gradflow_oracle as orc  # noqa: E402
from pygradflow_b200 import synth  # noqa: E402

KINDS = {"GradJac": "grad_jac", "KKT": "kkt", "Nominal": "nominal"}


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


@pytest.mark.parametrize("kind", list(KINDS))
@pytest.mark.parametrize("n,m,B", [(16, 8, 5), (24, 12, 3), (12, 0, 3)])
def test_create_scaling_weights_vs_oracle(kind, n, m, B):
    from pygradflow_b200.params import Params, ScalingType
    from pygradflow_b200.scale import create_scaling

    prob, refs, d = _batch(B, n, m)
    rng = np.random.default_rng(n + m)
    sp_ = rng.uniform(-2.0, 2.0, (B, n))
    sd_ = rng.standard_normal((B, m))
    sc = create_scaling(prob, Params(scaling_type=ScalingType[kind]), sp_, sd_)
    for b in range(B):
        ref = orc.create_scaling(refs[b], orc.OracleParams(scaling_type=KINDS[kind]), sp_[b], sd_[b])
        assert np.array_equal(sc.var_weights[b].cpu().numpy(), ref.var_weights), (kind, b)
        assert np.array_equal(sc.cons_weights[b].cpu().numpy(), ref.cons_weights), (kind, b)
        assert int(sc.obj_weight[b].item()) == ref.obj_weight


def test_scaled_callbacks_bitwise():
    """ScaledProblem (scale.py:153-231) through gf_ldexp == the oracle's scaled callbacks exactly, for the constant-
    derivative QP family and for the chained Rosenbrock family (Hessian rebuilt at every point)."""
    from pygradflow_b200.kernels import WorkList
    from pygradflow_b200.problem import BatchedRosenbrock
    from pygradflow_b200.scale import BatchedScaled, BatchedScaling

    f64 = dict(dtype=torch.float64, device="cuda")
    rng = np.random.default_rng(1)
    B, n, m = 4, 16, 8
    prob, refs, d = _batch(B, n, m)
    vw, cw, ow = rng.integers(-3, 4, (B, n)), rng.integers(-3, 4, (B, m)), rng.integers(-2, 3, B)
    sc = BatchedScaling(torch.as_tensor(vw, device="cuda"), torch.as_tensor(cw, device="cuda"), torch.as_tensor(ow))
    sp_ = BatchedScaled(prob, sc)
    x, y = rng.standard_normal((B, n)), rng.standard_normal((B, m))
    xt, yt = torch.as_tensor(x, **f64), torch.as_tensor(y, **f64)
    grad, cons, obj = torch.zeros((B, n), **f64), torch.zeros((B, m), **f64), torch.zeros((B,), **f64)
    w = WorkList.all(B)
    sp_.eval(xt, grad, cons, obj, w)
    J = sp_.jac(xt, torch.zeros((B, m, n), **f64), w)
    H = sp_.lag_hess(xt, yt, torch.zeros((B, n, n), **f64), w)
    for b in range(B):
        o = orc.ScaledProblem(refs[b], orc.Scaling(vw[b], cw[b], int(ow[b])))
        assert rel_err(grad[b].cpu().numpy(), o.obj_grad(x[b])) <= 1e-13      # the family's own H x + g: summation order
        assert rel_err(cons[b].cpu().numpy(), o.cons(x[b])) <= 1e-13
        assert abs(obj[b].item() - o.obj(x[b])) <= 1e-12 * max(1.0, abs(o.obj(x[b])))
        assert np.array_equal(J[b].cpu().numpy(), o.cons_jac(x[b]))
        assert np.array_equal(H[b].cpu().numpy(), o.lag_hess(x[b], y[b]))
        assert np.array_equal(sp_.var_lb[b].cpu().numpy(), o.var_lb) and np.array_equal(sp_.var_ub[b].cpu().numpy(), o.var_ub)
    # the scaling of the outputs alone is exact: scaled(x) == ldexp(unscaled(unscale(x)))
    g0, c0, o0 = torch.zeros((B, n), **f64), torch.zeros((B, m), **f64), torch.zeros((B,), **f64)
    prob.eval(sc.unscale_primal(xt).contiguous(), g0, c0, o0, w)
    assert torch.equal(grad, torch.ldexp(g0, -sc.var_weights + sc.obj_weight[:, None]))
    assert torch.equal(cons, torch.ldexp(c0, sc.cons_weights)) and torch.equal(obj, torch.ldexp(o0, sc.obj_weight))

    nr = 12
    dr = synth.rosenbrock_batch(range(B), nr)
    pr = BatchedRosenbrock(dr["a"], dr["b"], dr["lb"], dr["ub"])
    vwr, owr = rng.integers(-2, 3, (B, nr)), rng.integers(-2, 3, B)
    scr = BatchedScaling(torch.as_tensor(vwr, device="cuda"), torch.zeros((B, 0), dtype=torch.int32, device="cuda"),
                         torch.as_tensor(owr))
    spr = BatchedScaled(pr, scr)
    xr = torch.as_tensor(rng.uniform(-1, 1, (B, nr)), **f64)
    Hs = spr.lag_hess(xr, None, torch.zeros((B, nr, nr), **f64), w)
    H0 = pr.lag_hess(scr.unscale_primal(xr).contiguous(), None, torch.zeros((B, nr, nr), **f64), w)
    expo = scr.obj_weight[:, None, None] - scr.var_weights[:, :, None] - scr.var_weights[:, None, :]
    assert torch.equal(Hs, torch.ldexp(H0, expo))


@pytest.mark.parametrize("kind", ["KKT", "Nominal", "GradJac", "Custom"])
def test_scaled_solve_vs_oracle_and_reference(golden, kind):
    """solve_general with Params(scaling_type=...): per instance the oracle's (and, for the fixture instances, the
    reference's) status and solution.  The scaled random QPs open with a run of rejected steps on nearly singular
    systems, where the path depends on the rounding of the linear solve (see tests/test_oracle_golden.py), so
    iterates are compared at the solution and iteration counts only loosely."""
    from pygradflow_b200.params import Params, ScalingType
    from pygradflow_b200.problem import BatchedQP
    from pygradflow_b200.transform import solve_general

    g = golden("scaling")
    for (n, m, k) in [(16, 8, 0), (24, 12, 1)]:
        key = f"gqp_n{n}_m{m}_k{k}/{kind}"
        d = synth.general_qp_batch([k], n, m)
        prob = BatchedQP(d["H"], d["A"], d["g"], d["b"], d["lb"], d["ub"])
        kw = dict(scaling_type=ScalingType[kind], scaling_primal=g[f"{key}/scaling_primal"],
                  scaling_dual=g[f"{key}/scaling_dual"])
        okw = dict(scaling_type=kind.lower() if kind != "GradJac" else "grad_jac", scaling_primal=kw["scaling_primal"],
                   scaling_dual=kw["scaling_dual"])
        if kind == "Custom":
            kw["scaling"] = (g[f"{key}/var_weights"], g[f"{key}/cons_weights"], g[f"{key}/obj_weight"])
            okw["scaling"] = orc.Scaling(g[f"{key}/var_weights"], g[f"{key}/cons_weights"], int(g[f"{key}/obj_weight"]))
        res = solve_general(prob, d["cons_lb"], d["cons_ub"], Params(**kw), d["x0"], d["y0"])
        assert np.array_equal(res.scaling.var_weights[0].cpu().numpy(), g[f"{key}/var_weights"])
        assert np.array_equal(res.scaling.cons_weights[0].cpu().numpy(), g[f"{key}/cons_weights"])
        ref = orc.solve_general(orc.GeneralQP(d["H"][0], d["A"][0], d["g"][0], d["b"][0], d["lb"][0], d["ub"][0],
                                              d["cons_lb"][0], d["cons_ub"][0]), orc.OracleParams(**okw), d["x0"][0],
                                d["y0"][0])
        assert int(res.status[0].item()) == ref.status == int(g[f"{key}/status"]) == 1
        assert rel_err(res.x[0].cpu().numpy(), ref.x) <= 1e-5
        assert rel_err(res.x[0].cpu().numpy(), g[f"{key}/x"]) <= 1e-5
        assert rel_err(res.y[0].cpu().numpy(), g[f"{key}/y"]) <= 1e-3
        if kind in ("KKT", "Nominal"):  # the other two wander for 50-180 iterations on their rounding-dependent path
            assert abs(int(res.iterations[0].item()) - int(g[f"{key}/iterations"])) <= max(4, int(g[f"{key}/iterations"]) // 3)


def test_zero_scaling_is_identity():
    """Scaling.zero (scale.py:61-65; test_scale.py:22-48): the scaled solve is the unscaled one, bit for bit."""
    from pygradflow_b200.params import Params, ScalingType
    from pygradflow_b200.transform import solve_general

    B, n, m = 4, 16, 8
    prob, _, d = _batch(B, n, m)
    base = solve_general(prob, d["cons_lb"], d["cons_ub"], Params(), d["x0"], d["y0"])
    zero = (np.zeros((B, n), dtype=np.int64), np.zeros((B, m), dtype=np.int64))
    res = solve_general(prob, d["cons_lb"], d["cons_ub"], Params(scaling_type=ScalingType.Custom, scaling=zero),
                        d["x0"], d["y0"])
    assert torch.equal(res.iterations, base.iterations) and torch.equal(res.status, base.status)
    assert torch.equal(res.x, base.x) and torch.equal(res.y, base.y)
