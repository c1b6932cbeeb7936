"""The reference's iterative LinearSolvers (linear_solver/gmres_solver.py:7-35, minres_solver.py:6-24) as batched CUDA
kernels (gf_gmres_solve / gf_minres_solve): linear solves against the oracle (which calls the same scipy routines as the
reference) and against solutions of the REAL reference (tests/golden/iterative.npz) -- identical numbers of matrix-vector
products, solutions within 1e-9; `trans`, `initial_sol`, the failure code; whole solves with
Params.linear_solver_type = GMRES / MINRES against the reference's traces."""

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from helpers import rel_err  # noqa: E402
from oracle import gradflow_oracle as orc  # noqa: E402
from pygradflow_b200 import synth  # noqa: E402
from test_oracle_golden import ITER_MATS, check_iterative_trace  # noqa: E402


@pytest.fixture(scope="module", autouse=True)
def _gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def _check_solution(kind, mat, rhs, sol, ref, tag):
    """GMRES keeps an orthonormal basis (modified Gram-Schmidt), so two implementations that stop after the same number
    of products agree to rounding.  MINRES runs on the three-term Lanczos recurrence, which loses orthogonality after a
    few dozen steps: at equal product counts the iterates agree only to the accuracy of the solve itself (rtol = 1e-5),
    so there the check is the reference's own acceptance test (tests/pygradflow/test_linear_solver.py:94-136: the
    residual) plus agreement at the solve tolerance."""
    cond = np.linalg.cond(mat)
    if kind == "gmres":
        assert rel_err(sol, ref) <= 1e-9 * max(1.0, cond * 1e-3), (tag, rel_err(sol, ref))
        return
    r_sol, r_ref = np.linalg.norm(mat @ sol - rhs), np.linalg.norm(mat @ ref - rhs)
    assert r_sol <= 4.0 * r_ref + 1e-12 * np.linalg.norm(rhs), (tag, r_sol, r_ref)
    assert rel_err(sol, ref) <= 1e-5 * max(1.0, cond), (tag, rel_err(sol, ref))


@pytest.mark.parametrize("name", ITER_MATS)
def test_linear_solver_plugin_vs_reference(golden, name):
    """B200LinearSolver(method="gmres" / "minres") on the reference's fixtures: solve, solve(trans=True),
    solve(initial_sol=...) against the reference's own results."""
    from pygradflow_b200.plugin import B200LinearSolver, LinearSolverError, set_error_types

    set_error_types(None, None)
    g = golden("iterative")
    mat, rhs, x0 = g[f"{name}/mat"], g[f"{name}/rhs"], g[f"{name}/x0"]
    sym = bool(np.array_equal(mat, mat.T))
    for kind in ("gmres", "minres"):
        if kind == "minres" and not sym:
            continue
        s = B200LinearSolver(mat, symmetric=sym, method=kind)
        o = orc.make_linear_solver(mat, kind, symmetric=sym)
        cases = [("sol", {}), ("sol_x0", dict(initial_sol=lambda: x0.copy()))]
        if kind == "gmres":
            cases.append(("sol_trans", dict(trans=True)))
        for key, kw in cases:
            if bool(g[f"{name}/{kind}/{key}_failed"]):
                try:  # the failure is a non-convergence at the iteration limit: the same verdict is required
                    s.solve(rhs, **kw)
                    raise AssertionError((name, kind, key, "reference fails, CUDA path converged"))
                except Exception as err:
                    assert type(err).__name__ == "LinearSolverError", err
                continue
            sol = s.solve(rhs, **kw)
            ref = g[f"{name}/{kind}/{key}"]
            o.solve(rhs, **kw)
            assert int(s.iters.item()) == o.matvecs, (name, kind, key, int(s.iters.item()), o.matvecs)
            _check_solution(kind, mat, rhs, sol, ref, (name, key))


@pytest.mark.parametrize("kind", ["gmres", "minres"])
def test_batched_ragged_vs_oracle(kind):
    """A ragged batch (orders 1 ... 150 in one launch, work list with a device count): every instance against the
    oracle -- same product counts, solutions within 1e-9 (scaled by the conditioning)."""
    from pygradflow_b200 import kernels as K
    from pygradflow_b200.kernels import WorkList

    orders = [1, 2, 5, 12, 33, 64, 65, 100, 150, 48, 20, 7]
    B, ld = len(orders), max(orders)
    Km = np.zeros((B, ld, ld))
    R = np.zeros((B, ld))
    for b, N in enumerate(orders):
        if N >= 4:
            Kb, r, _ = synth.kkt_instance(N)
        else:
            rng = np.random.default_rng(N)
            M = rng.normal(size=(N, N))
            Kb, r = M @ M.T + np.eye(N), rng.normal(size=N)
        Km[b, :N, :N], R[b, :N] = Kb, r
    dev = "cuda"
    Kd = torch.as_tensor(Km, device=dev)
    rhs = torch.as_tensor(R, device=dev)
    Nvec = torch.as_tensor(orders, dtype=torch.int32, device=dev)
    info = torch.full((B,), -7, dtype=torch.int32, device=dev)
    iters = torch.zeros((B,), dtype=torch.int32, device=dev)
    rows = K.krylov_scratch_rows(kind == "minres")
    scratch = torch.zeros((B, rows, ld), dtype=torch.float64, device=dev)
    skip = 3  # instance left out by the work list keeps its rhs
    lst = torch.as_tensor([b for b in range(B) if b != skip] + [skip], dtype=torch.int32, device=dev)
    work = WorkList(lst, torch.as_tensor([B - 1], dtype=torch.int32, device=dev), B)
    if kind == "gmres":
        K.gmres_solve(Kd, ld, Nvec, rhs, None, None, False, scratch, info, iters, work)
    else:
        K.minres_solve(Kd, ld, Nvec, rhs, None, scratch, info, iters, work)
    sol, info, iters = rhs.cpu().numpy(), info.cpu().numpy(), iters.cpu().numpy()
    assert np.array_equal(sol[skip], R[skip]) and info[skip] == -7
    for b, N in enumerate(orders):
        if b == skip:
            continue
        o = orc.make_linear_solver(Km[b, :N, :N], kind, symmetric=True)
        try:
            ref = o.solve(R[b, :N])
        except orc.LinearSolverError:
            assert info[b] != 0, (b, N)
            continue
        assert info[b] == 0, (b, N, info[b])
        assert iters[b] == o.matvecs, (b, N, iters[b], o.matvecs)
        _check_solution(kind, Km[b, :N, :N], R[b, :N], sol[b, :N], ref, (b, N))


def test_gmres_failure_code_and_rejection():
    """A matrix GMRES(20) cannot solve within n restarts (a cyclic shift: the residual does not move until the last
    Krylov vector): info = n, the plug-in raises LinearSolverError like gmres_solver.py:32-33."""
    from pygradflow_b200.plugin import B200LinearSolver, set_error_types

    set_error_types(None, None)
    N = 60
    P = np.roll(np.eye(N), 1, axis=0)
    rhs = np.zeros(N)
    rhs[0] = 1.0
    with pytest.raises(orc.LinearSolverError):
        orc.make_linear_solver(P, "gmres", symmetric=False).solve(rhs)
    s = B200LinearSolver(P, symmetric=False, method="gmres")
    with pytest.raises(Exception) as ei:
        s.solve(rhs)
    assert type(ei.value).__name__ == "LinearSolverError" and f"error code {N}" in str(ei.value)


@pytest.mark.parametrize("newton", ["Simplified", "Full"])
@pytest.mark.parametrize("lin,form", [("GMRES", "Symmetric"), ("GMRES", "Asymmetric"), ("GMRES", "Extended"),
                                      ("MINRES", "Symmetric")])
def test_batched_solves_vs_reference(golden, lin, form, newton):
    """BatchedSolver with Params.linear_solver_type = GMRES / MINRES against Solver.solve of the real reference."""
    from pygradflow_b200.params import LinearSolverType, NewtonType, Params, StepSolverType
    from pygradflow_b200.problem import BatchedQP
    from pygradflow_b200.solver import BatchedSolver

    g = golden("iterative")
    for (n, m, k) in [(16, 8, 0), (32, 16, 2), (24, 0, 5)]:
        key = f"{lin}/{form}/{newton}/qp_n{n}_m{m}_k{k}"
        d = synth.qp_batch([k], n, m)
        prob = BatchedQP(d["H"], d["A"] if m else None, d["g"], d["b"] if m else None, d["lb"], d["ub"])
        params = Params(linear_solver_type=LinearSolverType[lin], step_solver_type=StepSolverType[form],
                        newton_type=NewtonType[newton], iteration_limit=400)
        solver = BatchedSolver(prob, params, use_graph=False)
        assert solver.engine.linear.name == lin
        trace = []

        def rec(outer, s):
            ph = int(s.phase[0].item())
            src = s.mid if ph == 2 else s.fin
            trace.append(dict(accept=ph in (2, 3), x=src[0][0].cpu().numpy().copy()))

        res = solver.solve(d["x0"], d["y0"] if m else None, on_iteration=rec)
        if bool(g[f"{key}/failed"]):
            assert int(res.status[0].item()) == 6, key  # GF_STATUS_LAMB_MAX: solver.py:323-326 raises
            continue

        class R:
            pass

        r = R()
        r.status, r.iterations = int(res.status[0].item()), int(res.iterations[0].item())
        r.x, r.trace = res.x[0].cpu().numpy(), trace
        check_iterative_trace(r, g, key)


def test_batched_gmres_matches_oracle_per_instance():
    """A batch of 6 QPs through GMRES on the Asymmetric formulation (the one that hands GMRES a start vector):
    status and optimum per instance against the oracle."""
    from pygradflow_b200.params import LinearSolverType, Params, StepSolverType
    from pygradflow_b200.problem import BatchedQP
    from pygradflow_b200.solver import BatchedSolver

    B, n, m = 6, 20, 8
    d = synth.qp_batch(range(B), n, m)
    prob = BatchedQP(d["H"], d["A"], d["g"], d["b"], d["lb"], d["ub"])
    params = Params(linear_solver_type=LinearSolverType.GMRES, step_solver_type=StepSolverType.Asymmetric,
                    iteration_limit=400)
    res = BatchedSolver(prob, params).solve(d["x0"], d["y0"])
    for b in range(B):
        p = orc.DenseQP(d["H"][b], d["A"][b], d["g"][b], d["b"][b], d["lb"][b], d["ub"][b])
        ref = orc.Solver(p, orc.OracleParams(linear_solver="gmres", step_solver_type="asymmetric",
                                             iteration_limit=400)).solve(d["x0"][b], d["y0"][b])
        assert int(res.status[b].item()) == ref.status, b
        assert rel_err(res.x[b].cpu().numpy(), ref.x) <= 1e-5, b
