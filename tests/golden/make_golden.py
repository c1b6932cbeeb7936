"""Generate golden fixtures by running the REAL reference (chrhansk/pygradflow v0.5.24).

Run in the build container only (the reference lives at /root/reference and cannot travel):

    python tests/golden/make_golden.py

Writes tests/golden/*.npz.  The fixtures pin the CPU oracle (oracle/gradflow_oracle.py) and,
on the GPU box, the CUDA path.  Problem data comes from pygradflow_b200.synth (seeded), the
reference-side Problem classes below follow tests/pygradflow/qp.py:4-30 of the reference
(scipy.sparse matrices, so the arithmetic goes through scipy's sparsetools like the
reference's own tests).
"""

import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(HERE, "_stubs"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import scipy.sparse as sps  # noqa: E402

from pygradflow.callbacks import CallbackType  # noqa: E402
from pygradflow.implicit_func import ImplicitFunc, ScaledImplicitFunc  # noqa: E402
from pygradflow.iterate import Iterate  # noqa: E402
from pygradflow.linear_solver import linear_solver  # noqa: E402
from pygradflow.newton import newton_method  # noqa: E402
from pygradflow.params import LinearSolverType, NewtonType, Params  # noqa: E402
from pygradflow.problem import Problem  # noqa: E402
from pygradflow.solver import Solver  # noqa: E402
from pygradflow.status import SolverStatus  # noqa: E402
from pygradflow.step.solver.symmetric_step_solver import SymmetricStepSolver  # noqa: E402

from pygradflow_b200 import synth  # noqa: E402


# ---------------------------------------------------------------- reference-side problems
class RefQP(Problem):
    def __init__(self, d):
        self.H = sps.csc_matrix(d["H"])
        self.A = sps.csr_matrix(d["A"])
        self.g = d["g"]
        self.b = d["b"]
        super().__init__(d["lb"], d["ub"], num_cons=d["A"].shape[0])

    def obj(self, x):
        return 0.5 * x @ (self.H @ x) + self.g @ x

    def obj_grad(self, x):
        return self.H @ x + self.g

    def cons(self, x):
        return self.A @ x + self.b

    def cons_jac(self, x):
        return self.A

    def lag_hess(self, x, _):
        return self.H


class RefChainedRosenbrock(Problem):
    def __init__(self, d):
        self.a = d["a"]
        self.b = d["b"]
        super().__init__(d["lb"], d["ub"])

    def obj(self, x):
        r = x[1:] - x[:-1] ** 2
        return float(np.sum(self.b * r * r + (self.a - x[:-1]) ** 2))

    def obj_grad(self, x):
        r = x[1:] - x[:-1] ** 2
        g = np.zeros_like(x)
        g[:-1] += -4.0 * self.b * r * x[:-1] - 2.0 * (self.a - x[:-1])
        g[1:] += 2.0 * self.b * r
        return g

    def cons(self, x):
        return np.array([])

    def cons_jac(self, x):
        return sps.coo_matrix(np.zeros((0, x.shape[0])))

    def lag_hess(self, x, _):
        n = x.shape[0]
        r = x[1:] - x[:-1] ** 2
        main = np.zeros(n)
        main[:-1] += 8.0 * self.b * x[:-1] ** 2 - 4.0 * self.b * r + 2.0
        main[1:] += 2.0 * self.b
        off = -4.0 * self.b * x[:-1]
        return sps.diags([off, main, off], [-1, 0, 1], format="csc")


class RefGeneralQP(Problem):
    """QP with general constraint bounds; Solver wraps it in the reference's ConstrainedProblem (slack transform)."""

    def __init__(self, d):
        self.H = sps.csc_matrix(d["H"])
        self.A = sps.csr_matrix(d["A"])
        self.g = d["g"]
        self.b = d["b"]
        super().__init__(d["lb"], d["ub"], cons_lb=d["cons_lb"], cons_ub=d["cons_ub"])

    def obj(self, x):
        return 0.5 * x @ (self.H @ x) + self.g @ x

    def obj_grad(self, x):
        return self.H @ x + self.g

    def cons(self, x):
        return self.A @ x + self.b

    def cons_jac(self, x):
        return self.A

    def lag_hess(self, x, _):
        return self.H


class RefOCP(Problem):
    """cfg4 family as a reference Problem: the formulas of the oracle's OCP class (identical arithmetic), returned
    as scipy.sparse matrices like the reference's own fixtures."""

    def __init__(self, d):
        from oracle import gradflow_oracle as orc

        self.p = orc.OCP(d["A"], d["B"], d["Q"], d["R"], d["xinit"], d["umax"], d["h"])
        super().__init__(self.p.var_lb, self.p.var_ub, num_cons=self.p.num_cons)

    def obj(self, x):
        return self.p.obj(x)

    def obj_grad(self, x):
        return self.p.obj_grad(x)

    def cons(self, x):
        return self.p.cons(x)

    def cons_jac(self, x):
        return sps.csr_matrix(self.p.cons_jac(x))

    def lag_hess(self, x, y):
        return sps.csc_matrix(self.p.lag_hess(x, y))


STATUS_CODE = {s: s.value for s in SolverStatus}


def _load_ref_fixture(name):
    """Load a problem class of the reference's own tests (tests/pygradflow/<name>.py) by path."""
    import importlib.util

    spec = importlib.util.spec_from_file_location(f"_ref_fixture_{name}", f"/root/reference/tests/pygradflow/{name}.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def ref_rosenbrock():
    return _load_ref_fixture("rosenbrock").Rosenbrock()


def ref_tame():
    return _load_ref_fixture("tame").Tame()


def ref_hs71():
    return _load_ref_fixture("hs71").HS71()


def trace_solve(problem, params, x0, y0, keep_every=1):
    """Run Solver.solve with the ComputedStep callback (solver.py:331) as trace hook."""
    solver = Solver(problem, params)
    xs, ys, accepts = [], [], []

    def cb(iterate, next_iterate, accept):
        xs.append(next_iterate.x.copy())
        ys.append(next_iterate.y.copy())
        accepts.append(bool(accept))

    solver.callbacks.register(CallbackType.ComputedStep, cb)
    res = solver.solve(x0, y0)
    xs = np.array(xs).reshape(len(accepts), -1)
    ys = np.array(ys).reshape(len(accepts), -1)
    sel = np.arange(0, len(accepts), keep_every)
    return dict(
        x=np.asarray(res.x),
        y=np.asarray(res.y),
        d=np.asarray(res.d),
        status=np.int32(STATUS_CODE[res.status]),
        iterations=np.int32(res.iterations),
        accepted_steps=np.int32(res.num_accepted_steps),
        accepts=np.array(accepts, dtype=bool),
        trace_idx=sel.astype(np.int32),
        trace_x=xs[sel],
        trace_y=ys[sel],
    )


def flat(prefix, d):
    return {f"{prefix}/{k}": v for k, v in d.items()}


def params_for(newton="Simplified", **kw):
    return Params(validate_input=False, newton_type=NewtonType[newton], **kw)


# ---------------------------------------------------------------- fixtures
def golden_linear_solver():
    """tests/pygradflow/test_linear_solver.py:19-79 through the reference LUSolver."""
    base = np.array(
        [[2, 1, 0, 0, 0], [1, 4, 1, 0, 1], [0, 1, 3, 2, 0], [0, 0, 2, -1, 0], [0, 1, 0, 0, 2]], dtype=float
    )
    ev = np.linalg.eigvalsh(base)
    posdef = base + np.diag([-min(2.0 * ev.min(), 0.0)] * 5)
    negdef = base + np.diag([-max(2.0 * ev.max(), 0.0)] * 5)
    rhs = np.array([4.0, 17.0, 19.0, 2.0, 12.0])
    out = {"rhs": rhs}
    for name, mat in (("indef", base), ("posdef", posdef), ("negdef", negdef)):
        s = linear_solver(sps.csc_matrix(mat), LinearSolverType.LU, symmetric=True)
        out[f"{name}/mat"] = mat
        out[f"{name}/sol"] = s.solve(rhs)
        out[f"{name}/sol_trans"] = s.solve(rhs, trans=True)
        out[f"{name}/neg"] = np.int32((np.linalg.eigvalsh(mat) < 0).sum())
    # cfg5-style quasi-definite KKT matrices
    for N in (12, 48, 96, 200):
        K, r, m = synth.kkt_instance(N)
        s = linear_solver(sps.csc_matrix(K), LinearSolverType.LU, symmetric=True)
        out[f"kkt{N}/sol"] = s.solve(r)
        out[f"kkt{N}/sol_trans"] = s.solve(r, trans=True)
        out[f"kkt{N}/neg"] = np.int32(m)
    np.savez_compressed(os.path.join(HERE, "linear_solver.npz"), **out)


def record_newton_steps(problem, params, x0, y0, dt, rho, nsteps=2):
    """newton_method(...).step twice, recording the dense K / rhs / sol of the step solver."""
    iterate = Iterate(problem, params, x0, y0)
    method = newton_method(problem, params, iterate, dt, rho)
    out = {}
    cur = iterate
    for j in range(nsteps):
        res = method.step(cur)
        ss = method.step_solver
        out[f"s{j}/active"] = np.asarray(res.active_set, dtype=bool)
        out[f"s{j}/dx"] = res.dx
        out[f"s{j}/dy"] = res.dy
        out[f"s{j}/xn"] = res.iterate.x
        out[f"s{j}/yn"] = res.iterate.y
        out[f"s{j}/diff"] = np.float64(res.diff)
        if isinstance(ss, SymmetricStepSolver) and ss._deriv is not None:
            out[f"s{j}/K"] = ss.deriv.toarray()
        func = ImplicitFunc(problem, iterate, dt)
        out[f"s{j}/Fnorm_unscaled"] = np.float64(np.linalg.norm(func.value_at(res.iterate, rho)))
        sfunc = ScaledImplicitFunc(problem, iterate, dt)
        out[f"s{j}/F_scaled_next"] = sfunc.value_at(res.iterate, rho)
        cur = res.iterate
    return out


def golden_newton():
    out = {}
    # QP family, a few sizes / seeds, one Newton step pair per Newton type
    for (n, m, k, dt, rho) in [(16, 8, 0, 1.0, 1e-8), (64, 32, 1, 1.0, 1e-2), (64, 32, 2, 0.05, 1.0), (48, 0, 3, 0.5, 1e-8)]:
        d = synth.qp_instance(k, n, m)
        prob = RefQP(d)
        rng = np.random.default_rng(77 + k)
        x0 = np.clip(rng.uniform(-1.2, 1.2, n), d["lb"], d["ub"])
        y0 = 0.1 * rng.standard_normal(m)
        for newton in ("Simplified", "Full", "ActiveSet"):
            key = f"qp_n{n}_m{m}_k{k}/{newton}"
            rec = record_newton_steps(prob, params_for(newton), x0, y0, dt, rho)
            out.update(flat(key, rec))
            out[f"{key}/x0"] = x0
            out[f"{key}/y0"] = y0
            out[f"{key}/dt"] = np.float64(dt)
            out[f"{key}/rho"] = np.float64(rho)
    # chained Rosenbrock
    for (n, k, dt, rho) in [(8, 0, 0.1, 1e-8), (64, 1, 0.02, 1e-8)]:
        d = synth.rosenbrock_instance(k, n)
        prob = RefChainedRosenbrock(d)
        for newton in ("Simplified", "Full", "ActiveSet"):
            key = f"ros_n{n}_k{k}/{newton}"
            rec = record_newton_steps(prob, params_for(newton), d["x0"], d["y0"], dt, rho)
            out.update(flat(key, rec))
            out[f"{key}/dt"] = np.float64(dt)
            out[f"{key}/rho"] = np.float64(rho)
    np.savez_compressed(os.path.join(HERE, "newton_steps.npz"), **out)


def golden_globalized():
    """GlobalizedNewtonMethod.step (newton.py:242-304) on small QPs; may raise on failure."""
    out = {}
    for (n, m, k, dt, rho) in [(16, 8, 0, 0.05, 1e-2), (32, 16, 4, 0.01, 1.0), (24, 0, 5, 0.1, 1e-8)]:
        d = synth.qp_instance(k, n, m)
        prob = RefQP(d)
        rng = np.random.default_rng(99 + k)
        x0 = np.clip(rng.uniform(-1.0, 1.0, n), d["lb"], d["ub"])
        y0 = 0.1 * rng.standard_normal(m)
        key = f"qp_n{n}_m{m}_k{k}"
        params = params_for("Globalized")
        iterate = Iterate(prob, params, x0, y0)
        method = newton_method(prob, params, iterate, dt, rho)
        cur = iterate
        for j in range(2):
            try:
                res = method.step(cur)
            except Exception as e:  # line search failure (newton.py:294)
                out[f"{key}/s{j}/failed"] = np.bool_(True)
                break
            out[f"{key}/s{j}/failed"] = np.bool_(False)
            out[f"{key}/s{j}/dx"] = res.dx
            out[f"{key}/s{j}/dy"] = res.dy
            out[f"{key}/s{j}/xn"] = res.iterate.x
            out[f"{key}/s{j}/yn"] = res.iterate.y
            out[f"{key}/s{j}/active"] = np.asarray(res.active_set, dtype=bool)
            cur = res.iterate
        out[f"{key}/x0"] = x0
        out[f"{key}/y0"] = y0
        out[f"{key}/dt"] = np.float64(dt)
        out[f"{key}/rho"] = np.float64(rho)
    np.savez_compressed(os.path.join(HERE, "globalized.npz"), **out)


def golden_solves():
    out = {}
    # cfg1: docs/solve_rosenbrock.py (2-D Rosenbrock, defaults) -> 30 its / 25 accepted
    res = trace_solve(ref_rosenbrock(), Params(), None, None)
    out.update(flat("rosenbrock2d", res))
    res = trace_solve(ref_tame(), params_for("Simplified"), np.zeros(2), np.zeros(1))
    out.update(flat("tame", res))
    for newton in ("Simplified", "Full", "ActiveSet"):
        res = trace_solve(ref_hs71(), params_for(newton), np.array([1.0, 5.0, 5.0, 1.0, 0.0]), np.zeros(2))
        out.update(flat(f"hs71/{newton}", res))
    # cfg3-style QPs (small), full traces
    for (n, m, k) in [(16, 8, 0), (16, 8, 1), (32, 16, 2), (64, 32, 3), (64, 32, 4), (48, 0, 5)]:
        d = synth.qp_instance(k, n, m)
        for newton in ("Simplified",) if n > 16 else ("Simplified", "Full", "ActiveSet"):
            res = trace_solve(RefQP(d), params_for(newton), d["x0"], d["y0"])
            out.update(flat(f"qp_n{n}_m{m}_k{k}/{newton}", res))
    # cfg2-style chained Rosenbrock: long runs, thin the stored trace
    for (n, k) in [(8, 0), (8, 1), (16, 2), (64, 1)]:
        d = synth.rosenbrock_instance(k, n)
        res = trace_solve(RefChainedRosenbrock(d), params_for("Simplified"), d["x0"], d["y0"], keep_every=10)
        out.update(flat(f"ros_n{n}_k{k}/Simplified", res))
    np.savez_compressed(os.path.join(HERE, "solves.npz"), **out)


def golden_full_size_qp():
    """One cfg3-size instance (n=512, m=256): status / counts / final point + thinned trace."""
    d = synth.qp_instance(0, 512, 256)
    res = trace_solve(RefQP(d), params_for("Simplified"), d["x0"], d["y0"], keep_every=8)
    np.savez_compressed(os.path.join(HERE, "qp512.npz"), **flat("qp_n512_m256_k0/Simplified", res))


def golden_constrained():
    """Slack transform (cons_problem.py, transform.py): general QPs and the reference's own HS71Constrained fixture;
    the recorded iterates are those of the transformed problem (x, slacks), the result is the restored one."""
    out = {}
    for (n, m, k) in [(16, 8, 0), (24, 12, 1)]:
        d = synth.general_qp_instance(k, n, m)
        res = trace_solve(RefGeneralQP(d), params_for("Simplified"), d["x0"], d["y0"])
        out.update(flat(f"gqp_n{n}_m{m}_k{k}/Simplified", res))
    hs = _load_ref_fixture("hs71_cons").HS71Constrained()
    res = trace_solve(hs, params_for("Simplified"), np.array([1.0, 5.0, 5.0, 1.0]), np.zeros(2))
    out.update(flat("hs71_cons/Simplified", res))
    np.savez_compressed(os.path.join(HERE, "constrained.npz"), **out)


class RefGeneralQPFresh(RefGeneralQP):
    """Returns NEW matrices from every callback.  The reference's ScaledProblem.cons_jac (scale.py:196-214) rescales
    `jac_orig.tocoo().data` in place, and csr.tocoo() shares the data array with the caller's matrix: a Problem that
    hands out its stored Jacobian (RefGeneralQP) has it rescaled again on every evaluation and the solve breaks down
    (lambda runs into lamb_max).  The reference's own fixtures build their matrices per call, as this class does."""

    def cons_jac(self, x):
        return self.A.copy()

    def lag_hess(self, x, _):
        return self.H.copy()


def golden_scaling():
    """Power-of-two scaling (scale.py) through the reference's Transformation: the weights create_scaling computes
    and Solver.solve with Params(scaling_type=...) on general QPs and on the reference's HS71Constrained fixture."""
    from pygradflow.params import ScalingType
    from pygradflow.scale import Scaling, create_scaling

    out = {}
    cases = []
    for (n, m, k) in [(16, 8, 0), (24, 12, 1)]:
        d = synth.general_qp_instance(k, n, m)
        cases.append((f"gqp_n{n}_m{m}_k{k}", lambda d=d: RefGeneralQPFresh(d), d["x0"], d["y0"],
                      np.linspace(0.3, 2.0, n), np.linspace(-1.0, 1.0, m)))
    hsmod = _load_ref_fixture("hs71_cons")
    cases.append(("hs71_cons", lambda: hsmod.HS71Constrained(), np.array([1.0, 5.0, 5.0, 1.0]), np.zeros(2),
                  np.array([1.0, 5.0, 5.0, 1.0]), np.array([0.5, -0.25])))
    for name, make, x0, y0, sp_, sd_ in cases:
        for kind in ("GradJac", "KKT", "Nominal", "Custom"):
            prob = make()
            kw = dict(scaling_type=ScalingType[kind], scaling_primal=sp_, scaling_dual=sd_)
            if kind == "Custom":
                rng = np.random.default_rng(len(name))
                kw["scaling"] = Scaling(rng.integers(-3, 4, prob.num_vars), rng.integers(-3, 4, prob.num_cons),
                                        int(rng.integers(-2, 3)))
            params = params_for("Simplified", **kw)
            sc = create_scaling(prob, params, sp_, sd_)
            try:
                res = trace_solve(prob, params, x0, y0, keep_every=4)
                res["failed"] = np.bool_(False)
            except Exception as err:  # solver.py:323-326: lambda ran into lamb_max
                res = dict(failed=np.bool_(True), error=np.str_(str(err)[:60]))
            res.update(var_weights=np.asarray(sc.var_weights, dtype=np.int64),
                       cons_weights=np.asarray(sc.cons_weights, dtype=np.int64), obj_weight=np.int64(sc.obj_weight),
                       scaling_primal=sp_, scaling_dual=sd_)
            out.update(flat(f"{name}/{kind}", res))
    np.savez_compressed(os.path.join(HERE, "scaling.npz"), **out)


def golden_controllers():
    """The other Newton-based step controllers (step_control.py:123-150): ResiduumRatio, Exact, Fixed."""
    from pygradflow.params import StepControlType

    out = {}
    for ctl in ("ResiduumRatio", "Exact", "Fixed"):
        kw = dict(step_control_type=StepControlType[ctl])
        if ctl == "Fixed":
            kw.update(iteration_limit=60)
        for (n, m, k) in [(16, 8, 0), (32, 16, 2)]:
            d = synth.qp_instance(k, n, m)
            res = trace_solve(RefQP(d), params_for("Simplified", **kw), d["x0"], d["y0"])
            out.update(flat(f"{ctl}/qp_n{n}_m{m}_k{k}", res))
        d = synth.rosenbrock_instance(0, 8)
        res = trace_solve(RefChainedRosenbrock(d), params_for("Simplified", **kw), d["x0"], d["y0"], keep_every=5)
        out.update(flat(f"{ctl}/ros_n8_k0", res))
    np.savez_compressed(os.path.join(HERE, "controllers.npz"), **out)


def golden_active_set_types():
    """tau-based active sets (newton_control.py:40-88, implicit_func.py:233-246)."""
    from pygradflow.params import ActiveSetType

    out = {}
    cases = [("Smallest", dict(active_set_type=ActiveSetType.SmallestActiveSet)),
             ("Largest", dict(active_set_type=ActiveSetType.LargestActiveSet)),
             ("Explicit", dict(active_set_type=ActiveSetType.Explicit, active_set_tau=0.3))]
    for name, kw in cases:
        for newton in ("Simplified", "Full"):
            for (n, m, k) in [(16, 8, 0), (32, 16, 2)]:
                d = synth.qp_instance(k, n, m)
                res = trace_solve(RefQP(d), params_for(newton, **kw), d["x0"], d["y0"])
                out.update(flat(f"{name}/{newton}/qp_n{n}_m{m}_k{k}", res))
    np.savez_compressed(os.path.join(HERE, "active_set_types.npz"), **out)


def golden_step_solvers():
    """The other step-solver formulations (step/solver/__init__.py:12-31): Asymmetric, Extended, Standard."""
    from pygradflow.params import StepSolverType

    out = {}
    for kind in ("Asymmetric", "Extended", "Standard"):
        kw = dict(step_solver_type=StepSolverType[kind])
        for newton in ("Simplified", "Full"):
            for (n, m, k) in [(16, 8, 0), (32, 16, 2), (24, 0, 5)]:
                d = synth.qp_instance(k, n, m)
                res = trace_solve(RefQP(d), params_for(newton, **kw), d["x0"], d["y0"])
                out.update(flat(f"{kind}/{newton}/qp_n{n}_m{m}_k{k}", res))
        d = synth.rosenbrock_instance(0, 8)
        res = trace_solve(RefChainedRosenbrock(d), params_for("Simplified", **kw), d["x0"], d["y0"], keep_every=5)
        out.update(flat(f"{kind}/Simplified/ros_n8_k0", res))
    np.savez_compressed(os.path.join(HERE, "step_solvers.npz"), **out)


def golden_penalty():
    """PenaltyUpdate.DualEquilibration (penalty.py:77-113) and Constant through Solver.solve."""
    from pygradflow.params import PenaltyUpdate

    out = {}
    for kind in ("DualEquilibration", "Constant"):
        for newton in ("Simplified", "Full"):
            for (n, m, k) in [(16, 8, 0), (32, 16, 2)]:
                d = synth.qp_instance(k, n, m)
                try:
                    res = trace_solve(RefQP(d), params_for(newton, penalty_update=PenaltyUpdate[kind], iteration_limit=300), d["x0"], d["y0"],
                                      keep_every=4)
                    res["failed"] = np.bool_(False)
                except Exception as err:  # solver.py:323-326: lambda ran into lamb_max
                    res = dict(failed=np.bool_(True), error=np.str_(str(err)[:60]))
                out.update(flat(f"{kind}/{newton}/qp_n{n}_m{m}_k{k}", res))
    np.savez_compressed(os.path.join(HERE, "penalty.npz"), **out)


def golden_rcond():
    """ConditionEstimator (step/cond_estimate.py) on the reference's LUSolver: the 5x5 fixtures of
    tests/pygradflow/test_linear_solver.py and cfg5-style KKT matrices."""
    from pygradflow.linear_solver.lu_solver import LUSolver
    from pygradflow.step.cond_estimate import ConditionEstimator

    g = np.load(os.path.join(HERE, "linear_solver.npz"))
    out = {}
    mats = {name: g[f"{name}/mat"] for name in ("indef", "posdef", "negdef")}
    for N in (12, 48, 96, 200):
        mats[f"kkt{N}"] = synth.kkt_instance(N)[0]
    for name, mat in mats.items():
        sm = sps.csc_matrix(mat)
        est = ConditionEstimator(sm, LUSolver(sm), Params())
        out[f"{name}/rcond"] = np.float64(est.estimate_rcond())
        out[f"{name}/its"] = np.int32(est._required_its())
    np.savez_compressed(os.path.join(HERE, "rcond.npz"), **out)


def golden_ocp():
    """cfg4-style discretised optimal-control problems (small): full traces of the real reference."""
    out = {}
    for (S, nx, nu, k) in [(6, 3, 2, 0), (16, 4, 3, 1)]:
        d = synth.ocp_instance(k, stages=S, nx=nx, nu=nu)
        for newton in ("Simplified", "Full"):
            res = trace_solve(RefOCP(d), params_for(newton), d["x0"], d["y0"], keep_every=4 if S > 6 else 1)
            out.update(flat(f"ocp_S{S}_nx{nx}_nu{nu}_k{k}/{newton}", res))
    np.savez_compressed(os.path.join(HERE, "ocp.npz"), **out)


def golden_ocp_full():
    """cfg4 at FULL size (S=128, nx=nu=8: n=2048, m=1024) through the real reference with scipy.sparse J / H:
    status, iteration counts, accept sequence, final iterate and every 25th candidate iterate."""
    out = {}
    for k in (0, 1, 2, 3):
        d = synth.ocp_instance(k)
        res = trace_solve(RefOCP(d), params_for("Simplified"), d["x0"], d["y0"], keep_every=25)
        out.update(flat(f"ocp_full_k{k}", res))
        print("ocp_full", k, int(res["status"]), int(res["iterations"]), flush=True)
    np.savez_compressed(os.path.join(HERE, "ocp_full.npz"), **out)


def golden_penalty_reject():
    """The penalty strategies that look at the candidate iterate (penalty.py:115-255): ParetoDecrease and the two
    filters (which can reject an accepted step, solver.py:357-378), whole Solver.solve traces incl. the rho sequence."""
    from pygradflow.params import PenaltyUpdate

    out = {}
    for kind in ("ParetoDecrease", "ObjectiveFilter", "LagrangianFilter"):
        for newton in ("Simplified", "Full"):
            cases = [("qp", (16, 8, 0)), ("qp", (32, 16, 2)), ("hs71", None), ("tame", None)]
            # random starts (x0 ~ U(-1,1), y0 ~ 2 N(0,1) from default_rng(500 + k)): the filters reject on these
            cases += [("qpr", (12, 5, k)) for k in (0, 2, 3, 10, 18, 32, 35)]
            for fam, spec in cases:
                if fam in ("qp", "qpr"):
                    n, m, k = spec
                    d = synth.qp_instance(k, n, m)
                    prob, x0, y0, name = RefQP(d), d["x0"], d["y0"], f"{fam}_n{n}_m{m}_k{k}"
                    if fam == "qpr":
                        rng = np.random.default_rng(500 + k)
                        x0, y0 = rng.uniform(-1, 1, n), rng.normal(size=m) * 2
                elif fam == "hs71":
                    prob, x0, y0, name = ref_hs71(), np.array([1.0, 5.0, 5.0, 1.0, 0.0]), np.zeros(2), "hs71"
                else:
                    prob, x0, y0, name = ref_tame(), np.zeros(2), np.zeros(1), "tame"
                rhos = []
                solver = Solver(prob, params_for(newton, penalty_update=PenaltyUpdate[kind], iteration_limit=300))
                xs, accepts = [], []

                def cb(iterate, next_iterate, accept):
                    xs.append(next_iterate.x.copy())
                    accepts.append(bool(accept))
                    rhos.append(float(solver.rho))

                solver.callbacks.register(CallbackType.ComputedStep, cb)
                try:
                    res = solver.solve(x0, y0)
                    r = dict(failed=np.bool_(False), x=np.asarray(res.x), y=np.asarray(res.y),
                             status=np.int32(STATUS_CODE[res.status]), iterations=np.int32(res.iterations),
                             accepted_steps=np.int32(res.num_accepted_steps), accepts=np.array(accepts, dtype=bool),
                             rhos=np.array(rhos), rho_final=np.float64(solver.rho), trace_x=np.array(xs))
                except Exception as err:
                    r = dict(failed=np.bool_(True), error=np.str_(repr(err)[:80]), accepts=np.array(accepts, dtype=bool),
                             rhos=np.array(rhos))
                out.update(flat(f"{kind}/{newton}/{name}", r))
                print(kind, newton, name, r.get("status"), r.get("iterations"), r.get("accepted_steps"),
                      r.get("rho_final"), r.get("error"), flush=True)
    np.savez_compressed(os.path.join(HERE, "penalty_reject.npz"), **out)


def golden_iterative():
    """The reference's iterative LinearSolvers (linear_solver/gmres_solver.py, minres_solver.py): solutions of its own
    5x5 fixtures and of cfg5-style KKT matrices (GMRES also transposed and with the start vector of
    AsymmetricStepSolver.initial_sol), and whole Solver.solve traces with Params.linear_solver_type = GMRES / MINRES."""
    from pygradflow.linear_solver import LinearSolverError
    from pygradflow.params import StepSolverType

    g = np.load(os.path.join(HERE, "linear_solver.npz"))
    rhs5 = g["rhs"]
    out = {}
    mats = {name: (g[f"{name}/mat"], rhs5) for name in ("indef", "posdef", "negdef")}
    for N in (12, 48, 96):
        K, r, _ = synth.kkt_instance(N)
        mats[f"kkt{N}"] = (K, r)
    rng = np.random.default_rng(77)
    for N in (10, 40):  # unsymmetric, well conditioned (exercises trans and x0)
        mats[f"unsym{N}"] = (np.eye(N) * 4.0 + rng.normal(size=(N, N)) / np.sqrt(N), rng.normal(size=N))
    for name, (mat, rhs) in mats.items():
        N = mat.shape[0]
        out[f"{name}/mat"] = mat
        out[f"{name}/rhs"] = rhs
        x0 = np.zeros(N)
        x0[::3] = rhs[::3]
        out[f"{name}/x0"] = x0
        sym = bool(np.array_equal(mat, mat.T))
        for kind, typ in (("gmres", LinearSolverType.GMRES), ("minres", LinearSolverType.MINRES)):
            if kind == "minres" and not sym:
                continue
            s = linear_solver(sps.csc_matrix(mat), typ, symmetric=sym)
            cases = [("sol", dict())]
            if kind == "gmres":
                cases += [("sol_trans", dict(trans=True)), ("sol_x0", dict(initial_sol=lambda x0=x0: x0.copy()))]
            else:
                cases += [("sol_x0", dict(initial_sol=lambda x0=x0: x0.copy()))]
            for key, kw in cases:
                try:
                    out[f"{name}/{kind}/{key}"] = s.solve(rhs, **kw)
                    out[f"{name}/{kind}/{key}_failed"] = np.bool_(False)
                except LinearSolverError:
                    out[f"{name}/{kind}/{key}_failed"] = np.bool_(True)
    for lin, forms in (("GMRES", ("Symmetric", "Asymmetric", "Extended")), ("MINRES", ("Symmetric",))):
        for form in forms:
            kw = dict(linear_solver_type=LinearSolverType[lin], step_solver_type=StepSolverType[form], iteration_limit=400)
            for newton in ("Simplified", "Full"):
                for (n, m, k) in [(16, 8, 0), (32, 16, 2), (24, 0, 5)]:
                    d = synth.qp_instance(k, n, m)
                    try:
                        res = trace_solve(RefQP(d), params_for(newton, **kw), d["x0"], d["y0"], keep_every=2)
                        res["failed"] = np.bool_(False)
                    except Exception as err:
                        res = dict(failed=np.bool_(True), error=np.str_(str(err)[:60]))
                    out.update(flat(f"{lin}/{form}/{newton}/qp_n{n}_m{m}_k{k}", res))
                    print(lin, form, newton, n, m, k, res.get("status"), res.get("iterations"), flush=True)
    np.savez_compressed(os.path.join(HERE, "iterative.npz"), **out)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        for name in sys.argv[1:]:
            globals()["golden_" + name]()
        sys.exit(0)
    golden_linear_solver()
    golden_newton()
    golden_globalized()
    golden_solves()
    golden_full_size_qp()
    golden_ocp()
    golden_constrained()
    golden_controllers()
    golden_active_set_types()
    golden_step_solvers()
    golden_scaling()
    golden_penalty()
    golden_rcond()
    golden_iterative()
    golden_penalty_reject()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))
