"""The full-order unsymmetric formulations of the Newton step (step/solver/__init__.py:12-31: Asymmetric, Extended,
Standard) as assembly modes of the batched engine: matrices and right-hand sides bit for bit against the oracle's restatement
of asymmetric_step_solver.py / extended_step_solver.py, whole solves per instance against the oracle and against
traces of the REAL reference (tests/golden/step_solvers.npz)."""

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


def _params(kind, newton="Simplified", **kw):
    from pygradflow_b200.params import NewtonType, Params, StepSolverType

    return Params(step_solver_type=StepSolverType[kind], newton_type=NewtonType[newton], **kw)


@pytest.mark.parametrize("kind", ["Asymmetric", "Extended", "Standard"])
@pytest.mark.parametrize("n,m,B", [(16, 8, 6), (40, 0, 5), (70, 33, 4)])
def test_full_order_system_bitwise(kind, n, m, B):
    """K and rhs of gf_kkt_assemble_full / gf_kkt_rhs_full == AsymmetricStepSolver / ExtendedStepSolver of the oracle
    (dense restatement of the reference's sparse bmat) exactly; the LU solution within 1e-10 of the oracle's."""
    from pygradflow_b200 import kernels as K
    from pygradflow_b200.engine import KKTEngine
    from pygradflow_b200.kernels import WorkList
    from pygradflow_b200.params import StepSolverType

    rng = np.random.default_rng(11 + n)
    d = synth.qp_batch(range(B), n, m)
    f64 = dict(dtype=torch.float64, device="cuda")
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), **f64)
    active = rng.uniform(size=(B, n)) < 0.3
    active[0] = False
    if B > 1:
        active[1, : n - 1] = True  # a single inactive variable
    F = rng.standard_normal((B, n + m))
    dt = 10.0 ** rng.uniform(-2, 1, B)
    rho = 10.0 ** rng.uniform(-8, 0, B)
    eng = KKTEngine(B, n, m, "cuda", formulation=StepSolverType[kind])
    eng.active.copy_(torch.as_tensor(active.astype(np.uint8)))
    w = WorkList.all(B)
    eng.update_active_set(w)
    H, J = t(d["H"]), (t(d["A"]) if m else None)
    eng.assemble(H, J, t(dt), t(rho), w)
    Kd = eng.K.cpu().numpy().copy()
    K.kkt_rhs_full(n, m, eng.perm, eng.nI, eng.active, t(F), t(dt), t(rho), eng.rhs, eng.form, w)
    rhs_d = eng.rhs.cpu().numpy().copy()
    eng.factor_assembled(H, J, t(dt), t(rho), w)
    eng.solve(eng.rhs, w)
    sol = eng.rhs.cpu().numpy()
    assert (eng.info.cpu().numpy() == 0).all()
    cls = {"Asymmetric": orc.AsymmetricStepSolver, "Extended": orc.ExtendedStepSolver,
           "Standard": orc.StandardStepSolver}[kind]
    for b in range(B):
        p = orc.DenseQP(d["H"][b], d["A"][b], d["g"][b], d["b"][b], d["lb"][b], d["ub"][b])
        it = orc.Iterate(p, orc.OracleParams(), d["x0"][b], d["y0"][b])
        s = cls(p, orc.OracleParams(), it, float(dt[b]), float(rho[b]))
        s.hess, s.jac, s.active_set = d["H"][b], d["A"][b].reshape(m, n), active[b]
        A = active[b]
        lamb = 1.0 / dt[b]
        fact = 1.0 / (1.0 + lamb * rho[b])
        if kind == "Standard":  # the matrix of the unscaled function on the handed-in H (= H_rho), rhs = F
            Kref, rref = s.matrix(), F[b]
        else:
            Kref, rref = s.full_system(dt[b] * F[b, :n][A], F[b, :n][~A], fact * F[b, n:])
        assert np.array_equal(Kd[b, : n + m, : n + m], Kref), (kind, b)
        assert np.array_equal(rhs_d[b, : n + m], rref), (kind, b)
        assert rel_err(sol[b, : n + m], np.linalg.solve(Kref, rref)) <= 1e-10 * max(1.0, np.linalg.cond(Kref) * 1e-4)


@pytest.mark.parametrize("newton", ["Simplified", "Full"])
@pytest.mark.parametrize("kind", ["Asymmetric", "Extended", "Standard"])
def test_formulations_qp_vs_oracle(kind, newton):
    from pygradflow_b200.problem import BatchedQP
    from pygradflow_b200.solver import BatchedSolver

    B, n, m = 8, 16, 8
    d = synth.qp_batch(range(B), n, m)
    prob = BatchedQP(d["H"], d["A"], d["g"], d["b"], d["lb"], d["ub"])
    solver = BatchedSolver(prob, _params(kind, newton))
    assert solver.engine.linear.name == "LU"
    res = solver.solve(d["x0"], d["y0"])
    for b in range(B):
        p = orc.DenseQP(d["H"][b], d["A"][b], d["g"][b], d["b"][b], d["lb"][b], d["ub"][b])
        ref = orc.Solver(p, orc.OracleParams(step_solver_type=kind.lower(), newton_type=newton.lower())).solve(
            d["x0"][b], d["y0"][b])
        assert int(res.status[b].item()) == ref.status, (b, kind)
        assert int(res.iterations[b].item()) == ref.iterations, (b, kind)
        assert int(res.accepted_steps[b].item()) == ref.accepted_steps, (b, kind)
        assert rel_err(res.x[b].cpu().numpy(), ref.x) <= 1e-8
        assert rel_err(res.y[b].cpu().numpy(), ref.y) <= 1e-8


@pytest.mark.parametrize("kind", ["Asymmetric", "Extended", "Standard"])
def test_formulations_golden_reference(golden, kind):
    """Against Solver.solve of the real reference with Params(step_solver_type=...)."""
    from pygradflow_b200.problem import BatchedQP, BatchedRosenbrock
    from pygradflow_b200.solver import BatchedSolver

    g = golden("step_solvers")
    for newton in ("Simplified", "Full"):
        for (n, m, k) in [(16, 8, 0), (32, 16, 2), (24, 0, 5)]:
            d = synth.qp_batch([k], n, m)
            prob = BatchedQP(d["H"], d["A"] if m else None, d["g"], d["b"] if m else None, d["lb"], d["ub"])
            res = BatchedSolver(prob, _params(kind, newton)).solve(d["x0"], d["y0"] if m else None)
            key = f"{kind}/{newton}/qp_n{n}_m{m}_k{k}"
            assert int(res.status[0].item()) == int(g[f"{key}/status"]), key
            assert int(res.iterations[0].item()) == int(g[f"{key}/iterations"]), key
            assert int(res.accepted_steps[0].item()) == int(g[f"{key}/accepted_steps"]), key
            assert rel_err(res.x[0].cpu().numpy(), g[f"{key}/x"]) <= 1e-8, key
    d = synth.rosenbrock_batch([0], 8)
    prob = BatchedRosenbrock(d["a"], d["b"], d["lb"], d["ub"])
    res = BatchedSolver(prob, _params(kind)).solve(d["x0"], None)
    key = f"{kind}/Simplified/ros_n8_k0"
    assert int(res.status[0].item()) == int(g[f"{key}/status"])
    assert int(res.iterations[0].item()) == int(g[f"{key}/iterations"])
    assert int(res.accepted_steps[0].item()) == int(g[f"{key}/accepted_steps"])
    assert rel_err(res.x[0].cpu().numpy(), g[f"{key}/x"]) <= 1e-7


def test_formulation_needs_lu():
    from pygradflow_b200.engine import KKTEngine
    from pygradflow_b200.params import LinearSolverType, StepSolverType

    with pytest.raises(ValueError):
        KKTEngine(2, 8, 4, "cuda", LinearSolverType.LDLT, formulation=StepSolverType.Asymmetric)


def test_formulations_agree_with_symmetric_at_cfg3_shape():
    """n = 192, m = 96 (order 288: the multi-launch LU with register-resident panels): the three formulations give the
    same iteration counts and iterates within 1e-8."""
    from pygradflow_b200.problem import BatchedQP
    from pygradflow_b200.solver import BatchedSolver

    B, n, m = 6, 192, 96
    d = synth.qp_batch(range(B), n, m)
    prob = BatchedQP(d["H"], d["A"], d["g"], d["b"], d["lb"], d["ub"])
    ref = BatchedSolver(prob, _params("Symmetric")).solve(d["x0"], d["y0"])
    for kind in ("Asymmetric", "Extended", "Standard"):
        res = BatchedSolver(prob, _params(kind)).solve(d["x0"], d["y0"])
        assert torch.equal(res.status, ref.status)
        assert torch.equal(res.iterations, ref.iterations), kind
        assert rel_err(res.x.cpu().numpy(), ref.x.cpu().numpy()) <= 1e-8
