"""Penalty strategies that look at the candidate iterate (penalty.py:115-255): ParetoDecrease (gf_pareto_update) and the
Objective / Lagrangian filters (gf_filter_update), which can veto an accepted step (solver.py:357-378).  Whole batched
solves against traces of the REAL reference (tests/golden/penalty_reject.npz: status, iteration and accepted-step
counts, the rho used in every iteration, the optimum) and per instance against the oracle."""

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from helpers import rel_err  # noqa: E402
from oracle import gradflow_oracle as orc  # noqa: E402
from pygradflow_b200 import synth  # noqa: E402
from test_oracle_golden import PENALTY_KIND  # noqa: E402

QPR = (0, 2, 3, 10, 18, 32, 35)


@pytest.fixture(scope="module", autouse=True)
def _gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def _qpr_batch():
    n, m = 12, 5
    d = synth.qp_batch(QPR, n, m)
    x0 = np.zeros((len(QPR), n))
    y0 = np.zeros((len(QPR), m))
    for i, k in enumerate(QPR):
        rng = np.random.default_rng(500 + k)
        x0[i], y0[i] = rng.uniform(-1, 1, n), rng.normal(size=m) * 2
    return d, x0, y0


def _filter_horizon(problem, kind, newton, x0, y0):
    """The filters compare (objective, violation) pairs of different iterates with <= (penalty.py:182-183).  Once two
    candidates' entries agree to rounding -- close to the optimum -- the verdict depends on the last bits of the
    objective evaluation (summation order), so from there on only status and optimum are comparable.  Returns the
    number of leading outer iterations whose verdicts are rounding-stable (all of them for ParetoDecrease)."""
    ref = orc.Solver(problem, orc.OracleParams(penalty_update=PENALTY_KIND[kind], newton_type=newton.lower(),
                                               iteration_limit=300)).solve(x0, y0, record=True)
    if kind == "ParetoDecrease":
        return ref, len(ref.trace)
    seen = []
    for i, t in enumerate(ref.trace):
        if not t["accept"]:
            continue
        it = orc.Iterate(problem, orc.OracleParams(), t["x"], t["y"])
        if kind == "ObjectiveFilter":
            e = (it.obj, it.cons_violation)
        else:
            dl = it.aug_lag_deriv_x(t["rho"])
            e = (float(dl @ dl + it.cons @ it.cons), float(np.linalg.norm(it.cons)))
        for f in seen:
            if any(abs(a - b) <= 1e-9 * max(1.0, abs(a)) for a, b in zip(e, f)):
                return ref, i
        seen.append(e)
    return ref, len(ref.trace)


def _solve(d, x0, y0, kind, newton, use_graph):
    from pygradflow_b200.params import NewtonType, Params, PenaltyUpdate
    from pygradflow_b200.problem import BatchedQP
    from pygradflow_b200.solver import BatchedSolver

    prob = BatchedQP(d["H"], d["A"], d["g"], d["b"], d["lb"], d["ub"])
    params = Params(penalty_update=PenaltyUpdate[kind], newton_type=NewtonType[newton], iteration_limit=300)
    solver = BatchedSolver(prob, params, use_graph=use_graph)
    B = prob.B
    rhos = [[] for _ in range(B)]
    hook = None
    if not use_graph:
        def hook(outer, s):
            st, ph, r = s.status.cpu().numpy(), s.phase.cpu().numpy(), s.rho.cpu().numpy()
            for b in range(B):
                if st[b] == 0 and ph[b] != 0:
                    rhos[b].append(float(r[b]))
    res = solver.solve(x0, y0, on_iteration=hook)
    return res, rhos


@pytest.mark.parametrize("newton", ["Simplified", "Full"])
@pytest.mark.parametrize("kind", list(PENALTY_KIND))
def test_penalty_strategies_vs_reference(golden, kind, newton):
    """The seven random-start QPs (the filters reject on them) as ONE batch, and two default-start QPs."""
    g = golden("penalty_reject")
    d, x0, y0 = _qpr_batch()
    res, rhos = _solve(d, x0, y0, kind, newton, use_graph=False)
    for i, k in enumerate(QPR):
        key = f"{kind}/{newton}/qpr_n12_m5_k{k}"
        p = orc.DenseQP(d["H"][i], d["A"][i], d["g"][i], d["b"][i], d["lb"][i], d["ub"][i])
        _, h = _filter_horizon(p, kind, newton, x0[i], y0[i])
        assert h >= 10, (key, h)
        assert int(res.status[i].item()) == int(g[f"{key}/status"]), key
        # the hook runs before the filter's veto / the commit: rho it sees is the one the iteration used
        gr = g[f"{key}/rhos"]
        assert np.allclose(np.array(rhos[i])[:h], gr[:h], rtol=1e-8, atol=0.0), key
        if h >= len(gr):
            assert int(res.iterations[i].item()) == int(g[f"{key}/iterations"]), key
            assert int(res.accepted_steps[i].item()) == int(g[f"{key}/accepted_steps"]), key
            assert rel_err(res.x[i].cpu().numpy(), g[f"{key}/x"]) <= 1e-8, key
        elif int(g[f"{key}/status"]) == 1:
            assert rel_err(res.x[i].cpu().numpy(), g[f"{key}/x"]) <= 1e-5, key
    for (n, m, k) in [(16, 8, 0), (32, 16, 2)]:
        key = f"{kind}/{newton}/qp_n{n}_m{m}_k{k}"
        dd = synth.qp_batch([k], n, m)
        r1, _ = _solve(dd, dd["x0"], dd["y0"], kind, newton, use_graph=False)
        assert int(r1.status[0].item()) == int(g[f"{key}/status"]), key
        assert int(r1.iterations[0].item()) == int(g[f"{key}/iterations"]), key
        assert abs(float(r1.rho[0].item()) - float(g[f"{key}/rho_final"])) <= 1e-9 * float(g[f"{key}/rho_final"]), key
        assert rel_err(r1.x[0].cpu().numpy(), g[f"{key}/x"]) <= 1e-8, key


@pytest.mark.parametrize("kind", list(PENALTY_KIND))
def test_penalty_strategies_graph_replay_vs_oracle(kind):
    """The same batch through the CUDA-graph replay of the outer loop: per instance identical to the oracle."""
    d, x0, y0 = _qpr_batch()
    res, _ = _solve(d, x0, y0, kind, "Simplified", use_graph=True)
    for i, k in enumerate(QPR):
        p = orc.DenseQP(d["H"][i], d["A"][i], d["g"][i], d["b"][i], d["lb"][i], d["ub"][i])
        ref, h = _filter_horizon(p, kind, "Simplified", x0[i], y0[i])
        assert int(res.status[i].item()) == ref.status, (kind, k)
        if h < len(ref.trace):  # verdicts past the rounding horizon of the filter: optimum only
            if ref.status == 1:
                assert rel_err(res.x[i].cpu().numpy(), ref.x) <= 1e-5, (kind, k)
            continue
        assert int(res.iterations[i].item()) == ref.iterations, (kind, k)
        assert int(res.accepted_steps[i].item()) == ref.accepted_steps, (kind, k)
        assert rel_err(res.x[i].cpu().numpy(), ref.x) <= 1e-8, (kind, k)
        if np.isfinite(ref.rho):
            assert abs(float(res.rho[i].item()) - ref.rho) <= 1e-9 * max(ref.rho, 1e-300), (kind, k)


def test_filter_capacity_overflow_raises():
    from pygradflow_b200.params import Params, PenaltyUpdate
    from pygradflow_b200.problem import BatchedQP
    from pygradflow_b200.solver import BatchedSolver

    d = synth.qp_batch([0, 1], 16, 8)
    prob = BatchedQP(d["H"], d["A"], d["g"], d["b"], d["lb"], d["ub"])
    params = Params(penalty_update=PenaltyUpdate.ObjectiveFilter, penalty_filter_capacity=1)
    with pytest.raises(RuntimeError, match="penalty filter capacity"):
        BatchedSolver(prob, params).solve(d["x0"], d["y0"])
