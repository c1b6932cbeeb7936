"""Round-2 parity items: the plug-in's full StepFunc surface (deriv_at, tau, projection_initial), the globalized
Newton method through the plug-in, Params.inertia_correction in the batched driver, the slack transform under the
Exact controller (work-list handling), and the full-size parity sweep of cfg2 / cfg3 / cfg4 against the CPU oracle."""

import json
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from helpers import noise_horizon, rel_err  # noqa: E402
from oracle import gradflow_oracle as orc  # noqa: E402
from pygradflow_b200 import synth  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module", autouse=True)
def _gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def _qp(n, m, k):
    d = synth.qp_instance(k, n, m)
    return orc.DenseQP(d["H"], d["A"], d["g"], d["b"], d["lb"], d["ub"]), d


# ------------------------------------------------------------------ plug-in StepFunc surface (implicit_func.py:202-294)
@pytest.mark.parametrize("n,m,k", [(12, 5, 0), (40, 16, 3), (24, 0, 1)])
def test_plugin_step_func_deriv_at_and_tau(n, m, k):
    from pygradflow_b200.plugin import B200StepFunc

    p, d = _qp(n, m, k)
    prm = orc.OracleParams()
    rng = np.random.default_rng(10 + k)
    it0 = orc.Iterate(p, prm, np.clip(rng.uniform(-1.2, 1.2, n), -1, 1), 0.1 * rng.standard_normal(m))
    it = orc.Iterate(p, prm, np.clip(it0.x + 0.3 * rng.standard_normal(n), -1, 1), it0.y + 0.05 * rng.standard_normal(m))
    dt, rho = 0.37, 0.02
    ours, ref = B200StepFunc(p, it0, dt), orc.ScaledImplicitFunc(p, it0, dt)
    A = ref.compute_active_set(it, rho)
    assert np.array_equal(ours.compute_active_set(it, rho), A)
    # F' of the scaled implicit function: copies of H_rho / J entries, lamb added on the diagonal -> bit-exact
    D = ours.deriv_at(it, rho)
    assert D.shape == (n + m, n + m)
    assert np.array_equal(D.toarray(), ref.deriv_at(it, rho))
    A2 = rng.uniform(size=n) < 0.4
    assert np.array_equal(ours.deriv_at(it, rho, A2).toarray(), ref.deriv_at(it, rho, A2))
    assert np.array_equal(ours.deriv(it.aug_lag_deriv_xy(), it.aug_lag_deriv_xx(rho), A2).toarray(), ref.deriv_at(it, rho, A2))
    for tau in (None, 0.05, 0.8, 2.5):
        assert rel_err(ours.projection_initial(it, rho, tau), ref.projection_initial(it, rho, tau)) <= 1e-13
        assert np.array_equal(ours.compute_active_set(it, rho, tau), ref.compute_active_set(it, rho, tau))
    pt = ref.projection_initial(it, rho)
    assert np.array_equal(ours.active_set_at_point(pt), ref.active_set_at_point(pt))
    assert np.array_equal(ours.project(pt, A), ref.project(pt, A))
    assert rel_err(ours.value_at(it, rho), ref.value_at(it, rho)) <= 1e-13
    assert rel_err(ours.value_at(it, rho, A2), ref.value_at(it, rho, A2)) <= 1e-13


@pytest.mark.parametrize("n,m,k", [(16, 8, 0), (32, 12, 4)])
def test_plugin_globalized_newton_method(n, m, k):
    """GlobalizedNewtonMethod (newton.py:218-304) drives the plug-in through func.deriv_at / value_at /
    compute_active_set and step_solver.solve(orig_iterate); same steps as with the oracle's own step solver."""
    from pygradflow_b200.plugin import B200StepSolver

    p, d = _qp(n, m, k)
    rng = np.random.default_rng(k)
    x0 = np.clip(rng.uniform(-1.3, 1.3, n), -1, 1)
    y0 = 0.2 * rng.standard_normal(m)
    res = {}
    for name, hook in (("ref", None), ("b200", B200StepSolver)):
        prm = orc.OracleParams(newton_type="globalized", step_solver=hook)
        it = orc.Iterate(p, prm, x0, y0)
        method = orc.newton_method(p, prm, it, 0.03, 1e-3)
        a = method.step(it)
        b = method.step(a.iterate)
        res[name] = (a, b, method.trials)
    for j in range(2):
        r, o = res["ref"][j], res["b200"][j]
        assert rel_err(o.iterate.x, r.iterate.x) <= 1e-10 and rel_err(o.iterate.y, r.iterate.y) <= 1e-10
        assert np.array_equal(o.active_set, r.active_set)
    assert res["ref"][2] == res["b200"][2]


def test_plugin_solver_with_tau_active_set():
    """Whole solve through the plug-in with ActiveSetType.SmallestActiveSet (tau handed to compute_active_set)."""
    from pygradflow_b200.plugin import B200StepSolver

    p, d = _qp(16, 8, 2)
    a = orc.Solver(p, orc.OracleParams(active_set_type="smallest")).solve(d["x0"], d["y0"], record=True)
    b = orc.Solver(p, orc.OracleParams(active_set_type="smallest", step_solver=B200StepSolver)).solve(d["x0"], d["y0"], record=True)
    assert a.status == b.status == 1
    assert [t["accept"] for t in a.trace][:10] == [t["accept"] for t in b.trace][:10]
    for i in range(min(8, len(a.trace), len(b.trace))):
        assert rel_err(b.trace[i]["x"], a.trace[i]["x"]) <= 1e-9
    assert rel_err(b.x, a.x) <= 1e-5


# ------------------------------------------------------------------ inertia correction (symmetric_step_solver.py:146-153)
def _nonconvex_qp_batch(B, n, m):
    """QPs whose Hessian has negative eigenvalues: with lambda = lamb_init the KKT matrix of some instances has the wrong
    inertia, so inertia_correction rejects steps (lambda doubles) where the plain solver would go on."""
    d = synth.qp_batch(range(B), n, m)
    for b in range(B):
        rng = np.random.default_rng(7000 + b)
        Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
        ev = np.concatenate([-rng.uniform(1.5, 4.0, n // 3), rng.uniform(0.2, 2.0, n - n // 3)])
        H = (Q * ev) @ Q.T
        d["H"][b] = 0.5 * (H + H.T)
    return d


@pytest.mark.parametrize("n,m,B", [(12, 4, 8), (80, 20, 6)])
def test_batched_inertia_correction_vs_oracle(n, m, B):
    """Target-problem style (tests/pygradflow/test_target_problem.py:46-61: Symmetric + inertia_correction on a
    non-convex problem).  The batched driver must reject exactly the steps the reference rejects."""
    from pygradflow_b200.params import Params
    from pygradflow_b200.problem import BatchedQP
    from test_gpu_newton import _batched_trace_solve

    d = _nonconvex_qp_batch(B, n, m)
    prob = BatchedQP(d["H"], d["A"], d["g"], d["b"], d["lb"], d["ub"])
    res, traces = _batched_trace_solve(prob, Params(inertia_correction=True, iteration_limit=60), d["x0"], d["y0"])
    rejected_for_inertia = 0
    for b in range(B):
        p = orc.DenseQP(d["H"][b], d["A"][b], d["g"][b], d["b"][b], d["lb"][b], d["ub"][b])
        ref = orc.Solver(p, orc.OracleParams(inertia_correction=True, iteration_limit=60, linear_solver="lapack")).solve(
            d["x0"][b], d["y0"][b], record=True)
        # strict up to the rounding-noise horizon of the QP family (helpers.noise_horizon), like the other trajectory tests
        h = min(noise_horizon(ref.trace), noise_horizon(traces[b]))
        k = min(len(traces[b]), len(ref.trace), h + 1)
        got = [(t["accept"], (not t["accept"]) and t["theta"] != t["theta"]) for t in traces[b][:k]]
        exp = [(t["accept"], t["newton_steps"] == 0) for t in ref.trace[:k]]
        assert got == exp, (b, got, exp)       # accepted / rejected by theta / rejected for the inertia: same sequence
        rejected_for_inertia += sum(1 for e in exp if e[1])
        for i in range(k):
            if not exp[i][1]:
                assert rel_err(traces[b][i]["x"], ref.trace[i]["x"]) <= 1e-9, (b, i)
        if h >= len(ref.trace):
            assert int(res.status[b].item()) == ref.status and int(res.iterations[b].item()) == ref.iterations
    assert rejected_for_inertia > 0  # the fixture must exercise the rejection path


def test_inertia_correction_needs_inertia():
    from pygradflow_b200.engine import KKTEngine
    from pygradflow_b200.params import LinearSolverType

    with pytest.raises(Exception, match="Inertia correction requested but not available"):
        KKTEngine(2, 8, 4, "cuda", LinearSolverType.LU, inertia_correction=True)
    assert KKTEngine(2, 8, 4, "cuda", LinearSolverType.Auto, inertia_correction=True).linear == LinearSolverType.LDLT


def test_plugin_inertia_correction_rejects():
    from pygradflow_b200 import plugin
    from pygradflow_b200.plugin import B200StepSolver

    plugin.set_error_types(orc.LinearSolverError, orc.StepSolverError)  # the host loop here is the oracle's
    d = _nonconvex_qp_batch(3, 12, 4)
    for b in range(3):
        p = orc.DenseQP(d["H"][b], d["A"][b], d["g"][b], d["b"][b], d["lb"][b], d["ub"][b])
        kw = dict(inertia_correction=True, iteration_limit=40)
        a = orc.Solver(p, orc.OracleParams(linear_solver="lapack", **kw)).solve(d["x0"][b], d["y0"][b], record=True)
        o = orc.Solver(p, orc.OracleParams(step_solver=B200StepSolver, **kw)).solve(d["x0"][b], d["y0"][b], record=True)
        k = min(len(a.trace), len(o.trace), 12)
        assert [t["accept"] for t in a.trace][:k] == [t["accept"] for t in o.trace][:k]
        assert [t["newton_steps"] for t in a.trace][:k] == [t["newton_steps"] for t in o.trace][:k]
    plugin.set_error_types(None, None)


# ------------------------------------------------------------------ slack transform + Exact controller (work lists)
@pytest.mark.parametrize("n,m,B", [(16, 8, 6)])
def test_solve_general_exact_controller_vs_oracle(n, m, B):
    """ExactController re-evaluates the mid / fin buffers with a shrinking work list; BatchedConstrained must leave the
    rows of instances outside the list untouched (their offsets / slacks are not applied twice)."""
    from pygradflow_b200.params import Params, StepControlType
    from pygradflow_b200.problem import BatchedQP
    from pygradflow_b200.transform import solve_general

    d = synth.general_qp_batch(range(B), n, m)
    prob = BatchedQP(d["H"], d["A"], d["g"], d["b"], d["lb"], d["ub"])
    res = solve_general(prob, d["cons_lb"], d["cons_ub"], Params(step_control_type=StepControlType.Exact), d["x0"], d["y0"])
    its = []
    for b in range(B):
        r = orc.GeneralQP(d["H"][b], d["A"][b], d["g"][b], d["b"][b], d["lb"][b], d["ub"][b], d["cons_lb"][b], d["cons_ub"][b])
        ref = orc.solve_general(r, orc.OracleParams(step_control_type="exact"), d["x0"][b], d["y0"][b])
        assert int(res.status[b].item()) == ref.status == 1
        assert rel_err(res.x[b].cpu().numpy(), ref.x) <= 1e-5
        assert int(res.iterations[b].item()) == ref.iterations
        its.append(ref.iterations)
    assert len(set(its)) > 1  # the instances finish at different times, so work lists shrink


def test_constrained_wrapper_respects_work_list():
    from pygradflow_b200.kernels import WorkList
    from pygradflow_b200.problem import BatchedQP
    from pygradflow_b200.transform import BatchedConstrained

    B, n, m = 5, 10, 6
    d = synth.general_qp_batch(range(B), n, m)
    cp = BatchedConstrained(BatchedQP(d["H"], d["A"], d["g"], d["b"], d["lb"], d["ub"]), d["cons_lb"], d["cons_ub"])
    f64 = dict(dtype=torch.float64, device="cuda")
    x = torch.as_tensor(np.random.default_rng(0).standard_normal((B, cp.n)), **f64)
    grad, cons, obj = torch.full((B, cp.n), 7.0, **f64), torch.full((B, m), 7.0, **f64), torch.full((B,), 7.0, **f64)
    wl = WorkList(torch.tensor([1, 3], dtype=torch.int32, device="cuda"), torch.tensor([2], dtype=torch.int32, device="cuda"), 2)
    cp.eval(x, grad, cons, obj, wl)
    g2, c2, o2 = torch.zeros_like(grad), torch.zeros_like(cons), torch.zeros_like(obj)
    cp.eval(x, g2, c2, o2, WorkList.all(B))
    for b in range(B):
        if b in (1, 3):
            assert torch.equal(grad[b], g2[b]) and torch.equal(cons[b], c2[b])
        else:
            assert bool((grad[b] == 7.0).all()) and bool((cons[b] == 7.0).all())


# ------------------------------------------------------------------ full-size parity sweep
def _run_sweep(cfg, count):
    import sys

    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import parity_sweep

    r = parity_sweep.sweep(cfg, count)
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, f"parity_sweep_cfg{cfg}.json"), "w") as f:
            json.dump(r, f, indent=1)
    return r["summary"], r["mismatches"]


@pytest.mark.parametrize("cfg,count", [(2, 128), (3, 128), (4, 8)])
def test_full_size_parity_sweep(cfg, count):
    """>= 128 instances of cfg2 (n=64) and cfg3 (n=512, m=256) and 8 of cfg4 (S=128: n=2048, m=1024) through
    BatchedSolver and through the oracle's Solver (splu) on the host: identical status everywhere; every instance
    that never reaches its rounding-noise horizon identical in iteration count / accept sequence / active sets /
    lambda trajectory; no instance differs BEFORE its horizon."""
    s, mism = _run_sweep(cfg, count)
    print(json.dumps(s))
    assert s["status_equal"] == count, mism[:4]
    assert s["diverged_before_horizon"] == 0, mism[:4]
    assert s["never_hit_horizon_identical"] == s["never_hit_horizon"], mism[:4]
    # whole solves: hundreds of adaptive steps, the final iterates of identical decision sequences agree to ~1e-9
    # (single Newton-KKT steps on identical inputs are checked at 1e-10 in test_gpu_newton.py / by bench.py)
    assert s["max_x_rel_identical"] <= 1e-8
