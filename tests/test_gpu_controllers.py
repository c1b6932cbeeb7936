"""The other Newton-based step controllers of the reference (step_control.py:123-150: ResiduumRatio, Exact, Fixed)
through the batched driver, per instance against the oracle and against traces of the REAL reference."""

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from helpers import rel_err  # noqa: E402
from oracle import gradflow_oracle as orc  # noqa: E402
from pygradflow_b200 import synth  # noqa: E402

NAMES = {"ResiduumRatio": "residuum_ratio", "Exact": "exact", "Fixed": "fixed"}


@pytest.fixture(scope="module", autouse=True)
def _gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def _params(ctl, newton="Simplified"):
    from pygradflow_b200.params import NewtonType, Params, StepControlType

    return Params(step_control_type=StepControlType[ctl], newton_type=NewtonType[newton],
                  iteration_limit=60 if ctl == "Fixed" else None)


@pytest.mark.parametrize("newton", ["Simplified", "Full"])
@pytest.mark.parametrize("ctl", ["ResiduumRatio", "Exact", "Fixed"])
def test_controllers_qp_vs_oracle(ctl, newton):
    from pygradflow_b200.problem import BatchedQP
    from pygradflow_b200.solver import BatchedSolver

    B, n, m = 8, 16, 8
    d = synth.qp_batch(range(B), n, m)
    prob = BatchedQP(d["H"], d["A"], d["g"], d["b"], d["lb"], d["ub"])
    res = BatchedSolver(prob, _params(ctl, newton)).solve(d["x0"], d["y0"])
    for b in range(B):
        p = orc.DenseQP(d["H"][b], d["A"][b], d["g"][b], d["b"][b], d["lb"][b], d["ub"][b])
        ref = orc.Solver(p, orc.OracleParams(step_control_type=NAMES[ctl], newton_type=newton.lower(),
                                             iteration_limit=60 if ctl == "Fixed" else None)).solve(d["x0"][b], d["y0"][b])
        assert int(res.status[b].item()) == ref.status, (b, ctl)
        assert int(res.iterations[b].item()) == ref.iterations, (b, ctl)
        assert int(res.accepted_steps[b].item()) == ref.accepted_steps, (b, ctl)
        assert rel_err(res.x[b].cpu().numpy(), ref.x) <= 1e-8
        assert rel_err(res.y[b].cpu().numpy(), ref.y) <= 1e-8


@pytest.mark.parametrize("ctl", ["ResiduumRatio", "Exact", "Fixed"])
def test_controllers_golden_reference(golden, ctl):
    from pygradflow_b200.problem import BatchedQP, BatchedRosenbrock
    from pygradflow_b200.solver import BatchedSolver

    g = golden("controllers")
    for (n, m, k) in [(16, 8, 0), (32, 16, 2)]:
        d = synth.qp_batch([k], n, m)
        prob = BatchedQP(d["H"], d["A"], d["g"], d["b"], d["lb"], d["ub"])
        res = BatchedSolver(prob, _params(ctl)).solve(d["x0"], d["y0"])
        key = f"{ctl}/qp_n{n}_m{m}_k{k}"
        assert int(res.status[0].item()) == int(g[f"{key}/status"])
        assert int(res.iterations[0].item()) == int(g[f"{key}/iterations"])
        assert int(res.accepted_steps[0].item()) == int(g[f"{key}/accepted_steps"])
        assert rel_err(res.x[0].cpu().numpy(), g[f"{key}/x"]) <= 1e-8
    d = synth.rosenbrock_batch([0], 8)
    prob = BatchedRosenbrock(d["a"], d["b"], d["lb"], d["ub"])
    res = BatchedSolver(prob, _params(ctl)).solve(d["x0"], None)
    key = f"{ctl}/ros_n8_k0"
    assert int(res.status[0].item()) == int(g[f"{key}/status"])
    assert int(res.iterations[0].item()) == int(g[f"{key}/iterations"])
    assert int(res.accepted_steps[0].item()) == int(g[f"{key}/accepted_steps"])
    assert rel_err(res.x[0].cpu().numpy(), g[f"{key}/x"]) <= 1e-7


@pytest.mark.parametrize("newton", ["Simplified", "Full"])
@pytest.mark.parametrize("kind", ["Smallest", "Largest", "Explicit"])
def test_tau_active_set_types_vs_oracle(kind, newton):
    """ActiveSetType Explicit / SmallestActiveSet / LargestActiveSet (newton_control.py:40-88): the active set is
    decided on the tau-variant of the projected point (implicit_func.py:237-244), per instance."""
    from pygradflow_b200.params import ActiveSetType, NewtonType, Params
    from pygradflow_b200.problem import BatchedQP
    from pygradflow_b200.solver import BatchedSolver

    B, n, m = 6, 16, 8
    d = synth.qp_batch(range(B), n, m)
    prob = BatchedQP(d["H"], d["A"], d["g"], d["b"], d["lb"], d["ub"])
    kw = {"Smallest": dict(active_set_type=ActiveSetType.SmallestActiveSet),
          "Largest": dict(active_set_type=ActiveSetType.LargestActiveSet),
          "Explicit": dict(active_set_type=ActiveSetType.Explicit, active_set_tau=0.3)}[kind]
    okw = {"Smallest": dict(active_set_type="smallest"), "Largest": dict(active_set_type="largest"),
           "Explicit": dict(active_set_type="explicit", active_set_tau=0.3)}[kind]
    res = BatchedSolver(prob, Params(newton_type=NewtonType[newton], **kw)).solve(d["x0"], d["y0"])
    for b in range(B):
        p = orc.DenseQP(d["H"][b], d["A"][b], d["g"][b], d["b"][b], d["lb"][b], d["ub"][b])
        ref = orc.Solver(p, orc.OracleParams(newton_type=newton.lower(), **okw)).solve(d["x0"][b], d["y0"][b])
        assert int(res.status[b].item()) == ref.status == 1
        if kind == "Largest":  # hundreds of iterations: past the rounding-noise horizon the paths differ slightly
            assert abs(int(res.iterations[b].item()) - ref.iterations) <= 0.1 * ref.iterations
            assert rel_err(res.x[b].cpu().numpy(), ref.x) <= 1e-4
        else:
            assert int(res.iterations[b].item()) == ref.iterations
            assert int(res.accepted_steps[b].item()) == ref.accepted_steps
            assert rel_err(res.x[b].cpu().numpy(), ref.x) <= 1e-8


def test_tau_active_set_types_golden_reference(golden):
    from pygradflow_b200.params import ActiveSetType, NewtonType, Params
    from pygradflow_b200.problem import BatchedQP
    from pygradflow_b200.solver import BatchedSolver

    g = golden("active_set_types")
    for name, kw in [("Smallest", dict(active_set_type=ActiveSetType.SmallestActiveSet)),
                     ("Explicit", dict(active_set_type=ActiveSetType.Explicit, active_set_tau=0.3))]:
        for newton in ("Simplified", "Full"):
            for (n, m, k) in [(16, 8, 0), (32, 16, 2)]:
                d = synth.qp_batch([k], n, m)
                prob = BatchedQP(d["H"], d["A"], d["g"], d["b"], d["lb"], d["ub"])
                res = BatchedSolver(prob, Params(newton_type=NewtonType[newton], **kw)).solve(d["x0"], d["y0"])
                key = f"{name}/{newton}/qp_n{n}_m{m}_k{k}"
                assert int(res.status[0].item()) == int(g[f"{key}/status"])
                assert int(res.iterations[0].item()) == int(g[f"{key}/iterations"])
                assert rel_err(res.x[0].cpu().numpy(), g[f"{key}/x"]) <= 1e-8


@pytest.mark.parametrize("newton", ["Simplified", "Full"])
@pytest.mark.parametrize("kind", ["DualEquilibration", "Constant"])
def test_penalty_strategies_vs_oracle(kind, newton):
    """PenaltyUpdate.DualEquilibration / Constant (penalty.py:36-113) in gf_commit: after exactly 10 outer iterations
    (rounding-stable stretch: rho reaches 1e7 within 20 iterations, see tests/test_oracle_golden.py::test_penalty_strategies) iterate, rho and lambda of every
    instance equal the oracle's; with the reference's limit of 300 the status (iteration limit / lamb_max) matches."""
    from pygradflow_b200.params import NewtonType, Params, PenaltyUpdate
    from pygradflow_b200.problem import BatchedQP
    from pygradflow_b200.solver import BatchedSolver

    B, n, m = 6, 16, 8
    d = synth.qp_batch(range(B), n, m)
    prob = BatchedQP(d["H"], d["A"], d["g"], d["b"], d["lb"], d["ub"])
    name = {"DualEquilibration": "dual_equilibration", "Constant": "constant"}[kind]
    for limit in (10, 300):
        res = BatchedSolver(prob, Params(penalty_update=PenaltyUpdate[kind], newton_type=NewtonType[newton],
                                         iteration_limit=limit)).solve(d["x0"], d["y0"])
        for b in range(B):
            p = orc.DenseQP(d["H"][b], d["A"][b], d["g"][b], d["b"][b], d["lb"][b], d["ub"][b])
            op = orc.OracleParams(penalty_update=name, newton_type=newton.lower(), iteration_limit=limit)
            if limit == 300 and kind == "DualEquilibration":
                # rho has grown until the KKT systems are nearly singular: which of the two ends an instance meets
                # (iteration limit, lamb_max = GF_STATUS_LAMB_MAX) depends on the rounding of the linear solve
                assert int(res.status[b].item()) in (2, 6), (kind, b)
                continue
            ref = orc.Solver(p, op).solve(d["x0"][b], d["y0"][b])
            assert int(res.status[b].item()) == ref.status, (kind, limit, b)
            if True:
                assert int(res.iterations[b].item()) == ref.iterations
                assert rel_err(res.x[b].cpu().numpy(), ref.x) <= 1e-8
                assert abs(res.rho[b].item() - ref.rho) <= 1e-10 * max(1.0, ref.rho)
                assert abs(res.lamb[b].item() - ref.lamb) <= 1e-6 * max(1.0, ref.lamb) + 1e-4 * ref.lamb
