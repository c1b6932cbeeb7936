"""gf_rosen_fused_solve: the whole Solver.solve of small chained-Rosenbrock instances inside one kernel (one warp per
instance, no lock-step) against the CPU oracle, the REAL reference's traces (tests/golden/solves.npz) and the lock-step
driver built from the stand-alone kernels: identical status, iteration and accepted-step counts, final active sets;
iterates within 1e-8."""

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


def _solve(d, fused=True, chunk=None, **kw):
    from pygradflow_b200.params import Params
    from pygradflow_b200.problem import BatchedRosenbrock
    from pygradflow_b200.solver import BatchedSolver

    prob = BatchedRosenbrock(d["a"], d["b"], d["lb"], d["ub"])
    s = BatchedSolver(prob, Params(fused=fused, **kw))
    if chunk is not None:
        s.FUSED_CHUNK = chunk
    res = s.solve(d["x0"], None)
    assert (getattr(s, "fused_launches", 0) > 0) == fused
    return s, res


@pytest.mark.parametrize("n,B", [(8, 8), (16, 6), (33, 5), (64, 24)])
def test_fused_vs_oracle_and_lockstep(n, B):
    d = synth.rosenbrock_batch(range(B), n)
    s, res = _solve(d)
    _, lock = _solve(d, fused=False)
    act = s.engine.active.cpu().numpy().astype(bool)
    all_pre = True
    for b in range(B):
        p = orc.ChainedRosenbrock(d["a"][b], d["b"][b], d["lb"][b], d["ub"][b])
        ref = orc.Solver(p, orc.OracleParams()).solve(d["x0"][b], d["y0"][b], record=True)
        assert int(res.status[b].item()) == ref.status, b
        if noise_horizon(ref.trace) < len(ref.trace):  # rounding-noise theta fed to the PI controller: optimum only
            assert rel_err(res.x[b].cpu().numpy(), ref.x) <= 1e-4, b
            all_pre = False
            continue
        assert int(res.iterations[b].item()) == ref.iterations, b
        assert int(res.accepted_steps[b].item()) == ref.accepted_steps, b
        assert int(res.iterations[b].item()) == int(lock.iterations[b].item()), b
        assert rel_err(res.x[b].cpu().numpy(), ref.x) <= 1e-8, b
        assert rel_err(res.x[b].cpu().numpy(), lock.x[b].cpu().numpy()) <= 1e-8, b
        last = [t["active"] for t in ref.trace if t["active"] is not None][-1]
        assert np.array_equal(act[b], last), b
        assert abs(res.lamb[b].item() - ref.lamb) <= 1e-3 * ref.lamb, b  # lambda carries the rounding of |d2| (tools/parity_sweep.py)
    if all_pre:
        assert res.newton_steps == lock.newton_steps


def test_fused_reference_traces(golden):
    """Against Solver.solve of the real reference: cfg1 (docs/solve_rosenbrock.py: 30 iterations, 25 accepted,
    x = [0.99999959 0.99999917]) and the cfg2-style instances of tests/golden/solves.npz."""
    g = golden("solves")
    d = dict(a=np.array([[1.0]]), b=np.array([[100.0]]), lb=np.full((1, 2), -np.inf), ub=np.full((1, 2), np.inf),
             x0=np.array([[0.0, 0.0]]))
    _, res = _solve(d)
    assert int(res.iterations[0].item()) == int(g["rosenbrock2d/iterations"]) == 30
    assert int(res.accepted_steps[0].item()) == int(g["rosenbrock2d/accepted_steps"]) == 25
    assert rel_err(res.x[0].cpu().numpy(), g["rosenbrock2d/x"]) <= 1e-10
    for (n, k) in [(8, 0), (8, 1), (16, 2), (64, 1)]:
        key = f"ros_n{n}_k{k}/Simplified"
        dd = synth.rosenbrock_batch([k], n)
        _, r = _solve(dd)
        assert int(r.status[0].item()) == int(g[f"{key}/status"]), key
        assert int(r.iterations[0].item()) == int(g[f"{key}/iterations"]), key
        assert int(r.accepted_steps[0].item()) == int(g[f"{key}/accepted_steps"]), key
        assert rel_err(r.x[0].cpu().numpy(), g[f"{key}/x"]) <= 1e-8, key


def test_fused_chunked_relaunch_and_iteration_limit():
    """The kernel's state survives a relaunch bit for bit (FUSED_CHUNK = 7 outer iterations per launch), and the
    iteration limit ends an instance with SolverStatus.IterationLimit exactly like the lock-step driver."""
    d = synth.rosenbrock_batch(range(10), 64)
    _, one = _solve(d)
    s, many = _solve(d, chunk=7)
    assert s.fused_launches > 10
    assert torch.equal(one.x, many.x) and torch.equal(one.iterations, many.iterations)
    assert torch.equal(one.lamb, many.lamb) and torch.equal(one.status, many.status)
    _, lim = _solve(d, iteration_limit=50)
    _, lock = _solve(d, fused=False, iteration_limit=50)
    assert (lim.status == 2).all() and (lim.iterations == 50).all()
    assert torch.equal(lim.accepted_steps, lock.accepted_steps)
    assert rel_err(lim.x.cpu().numpy(), lock.x.cpu().numpy()) <= 1e-8
