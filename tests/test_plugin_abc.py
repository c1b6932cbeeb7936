"""Structural check of the drop-in classes against the REAL reference ABCs (CPU, build container only).

The reference lives at /root/reference and cannot travel to the GPU box, so the plug-in can never run under the
reference's own ``Solver`` there; here, where it is importable, every public member of the reference's
``StepSolver`` (step/solver/step_solver.py:66-130), ``LinearSolver`` (linear_solver/linear_solver.py:18-31),
``StepFunc`` / ``ScaledImplicitFunc`` (implicit_func.py:12-99,202-294) and ``StepResult`` (step_solver.py:16-63) must
exist on the B200 class with the same parameter names, so that ``Params(step_solver=B200StepSolver)`` satisfies every
call site of newton.py / newton_control.py / the controllers.
"""
import inspect
import os
import sys

import pytest

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "pygradflow")), reason="reference not present")


@pytest.fixture(scope="module")
def ref():
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(here, "golden", "_stubs"))
    sys.path.insert(0, REF)
    try:
        import importlib

        imf = importlib.import_module("pygradflow.implicit_func")
        ls = importlib.import_module("pygradflow.linear_solver.linear_solver")
        ss = importlib.import_module("pygradflow.step.solver.step_solver")

        yield dict(StepSolver=ss.StepSolver, StepResult=ss.StepResult, LinearSolver=ls.LinearSolver,
                   LinearSolverError=ls.LinearSolverError, StepFunc=imf.StepFunc, Scaled=imf.ScaledImplicitFunc)
    finally:
        sys.path.remove(REF)


def _public(cls):
    return sorted(n for n in dir(cls) if not n.startswith("_"))


def _params(fn):
    fn = fn.fget if isinstance(fn, property) else fn
    return [p for p in inspect.signature(fn).parameters if p != "self"]


def _check(ref_cls, ours, skip=()):
    missing = [n for n in _public(ref_cls) if n not in skip and not hasattr(ours, n)]
    assert not missing, f"{ours.__name__} lacks {missing} of {ref_cls.__name__}"
    for n in _public(ref_cls):
        if n in skip:
            continue
        r, o = inspect.getattr_static(ref_cls, n), inspect.getattr_static(ours, n)
        assert isinstance(r, property) == isinstance(o, property), n
        if callable(r) or isinstance(r, property):
            rp, op = _params(r), _params(o)
            assert op[: len(rp)] == rp, (ours.__name__, n, rp, op)


def test_step_solver_covers_reference_abc(ref):
    from pygradflow_b200.plugin import B200StepSolver

    _check(ref["StepSolver"], B200StepSolver)
    for name in ref["StepSolver"].__abstractmethods__:
        assert hasattr(B200StepSolver, name)
    # constructor as the plug-in hook calls it: step_solver(problem, params, iterate, dt, rho) (step/solver/__init__.py:18-19)
    assert _params(B200StepSolver.__init__)[:5] == ["problem", "params", "orig_iterate", "dt", "rho"]


def test_linear_solver_covers_reference_abc(ref):
    from pygradflow_b200.plugin import B200LinearSolver

    _check(ref["LinearSolver"], B200LinearSolver)
    assert _params(B200LinearSolver.__init__)[:2] == ["matrix", "symmetric"]


def test_step_func_covers_reference(ref):
    from pygradflow_b200.plugin import B200StepFunc

    _check(ref["StepFunc"], B200StepFunc)
    _check(ref["Scaled"], B200StepFunc)


def test_step_result_covers_reference(ref):
    from pygradflow_b200.plugin import StepResult

    for n in ("iterate", "diff"):
        assert isinstance(inspect.getattr_static(StepResult, n), property)
    assert _params(StepResult.__init__)[:5] == _params(ref["StepResult"].__init__)[:5]


def test_plugin_raises_reference_exception_types(ref):
    from pygradflow.step.step_solver_error import StepSolverError

    from pygradflow_b200 import plugin

    lse, sse = plugin._reference_errors()
    assert lse is ref["LinearSolverError"] and sse is StepSolverError
