"""Pin the CPU oracle (oracle/gradflow_oracle.py) against fixtures produced by the real
reference (tests/golden/make_golden.py).  CPU only.

Tolerances: iterates and KKT solutions 1e-10 relative (inf-norm, relative to max(1, |.|));
active sets, accept sequences, iteration counts and status identical.
"""

import numpy as np
import pytest

from helpers import noise_horizon, rel_err
from oracle import gradflow_oracle as orc
from pygradflow_b200 import synth

RTOL = 1e-10


def make_qp(n, m, k):
    d = synth.qp_instance(k, n, m)
    return orc.DenseQP(d["H"], d["A"], d["g"], d["b"], d["lb"], d["ub"]), d


def make_ros(n, k):
    d = synth.rosenbrock_instance(k, n)
    return orc.ChainedRosenbrock(d["a"], d["b"], d["lb"], d["ub"]), d


NEWTON = {"Simplified": "simplified", "Full": "full", "ActiveSet": "active_set"}


# ------------------------------------------------------------------ linear solver
@pytest.mark.parametrize("kind", ["splu", "lapack"])
@pytest.mark.parametrize("name", ["indef", "posdef", "negdef"])
def test_linear_solver_fixtures(golden, name, kind):
    g = golden("linear_solver")
    mat, rhs = g[f"{name}/mat"], g["rhs"]
    s = orc.OracleLUSolver(mat, kind, symmetric=True)
    sol = s.solve(rhs)
    assert np.allclose(mat @ sol - rhs, 0.0)  # the reference's own criterion (test_linear_solver.py:101)
    assert rel_err(sol, g[f"{name}/sol"]) <= RTOL
    assert rel_err(s.solve(rhs, trans=True), g[f"{name}/sol_trans"]) <= RTOL
    assert s.num_neg_eigvals() == int(g[f"{name}/neg"])


@pytest.mark.parametrize("N", [12, 48, 96, 200])
def test_linear_solver_kkt(golden, N):
    g = golden("linear_solver")
    K, rhs, m = synth.kkt_instance(N)
    for kind in ("splu", "lapack"):
        s = orc.OracleLUSolver(K, kind, symmetric=True)
        assert rel_err(s.solve(rhs), g[f"kkt{N}/sol"]) <= RTOL
        assert rel_err(s.solve(rhs, trans=True), g[f"kkt{N}/sol_trans"]) <= RTOL
    assert s.num_neg_eigvals() == m == int(g[f"kkt{N}/neg"])


# ------------------------------------------------------------------ Newton steps
def _check_newton_steps(g, key, problem, x0, y0, newton):
    params = orc.OracleParams(newton_type=NEWTON[newton])
    dt, rho = float(g[f"{key}/dt"]), float(g[f"{key}/rho"])
    it = orc.Iterate(problem, params, x0, y0)
    method = orc.newton_method(problem, params, it, dt, rho)
    cur = it
    for j in range(2):
        res = method.step(cur)
        assert np.array_equal(res.active_set, g[f"{key}/s{j}/active"])
        assert rel_err(res.dx, g[f"{key}/s{j}/dx"]) <= RTOL
        assert rel_err(res.dy, g[f"{key}/s{j}/dy"]) <= RTOL
        assert rel_err(res.iterate.x, g[f"{key}/s{j}/xn"]) <= RTOL
        assert rel_err(res.iterate.y, g[f"{key}/s{j}/yn"]) <= RTOL
        assert abs(res.diff - float(g[f"{key}/s{j}/diff"])) <= RTOL * max(1.0, res.diff)
        if f"{key}/s{j}/K" in g.files:
            K = method.step_solver.K
            assert np.array_equal(K, g[f"{key}/s{j}/K"])  # pure copies: bit-exact
        f = orc.ImplicitFunc(problem, it, dt)
        fn = float(np.linalg.norm(f.value_at(res.iterate, rho)))
        assert abs(fn - float(g[f"{key}/s{j}/Fnorm_unscaled"])) <= 1e-9 * max(1.0, fn)
        sf = orc.ScaledImplicitFunc(problem, it, dt)
        assert rel_err(sf.value_at(res.iterate, rho), g[f"{key}/s{j}/F_scaled_next"]) <= 1e-9
        cur = res.iterate


@pytest.mark.parametrize("newton", list(NEWTON))
@pytest.mark.parametrize("n,m,k", [(16, 8, 0), (64, 32, 1), (64, 32, 2), (48, 0, 3)])
def test_newton_steps_qp(golden, n, m, k, newton):
    g = golden("newton_steps")
    key = f"qp_n{n}_m{m}_k{k}/{newton}"
    problem, _ = make_qp(n, m, k)
    _check_newton_steps(g, key, problem, g[f"{key}/x0"], g[f"{key}/y0"], newton)


@pytest.mark.parametrize("newton", list(NEWTON))
@pytest.mark.parametrize("n,k", [(8, 0), (64, 1)])
def test_newton_steps_rosenbrock(golden, n, k, newton):
    g = golden("newton_steps")
    key = f"ros_n{n}_k{k}/{newton}"
    problem, d = make_ros(n, k)
    _check_newton_steps(g, key, problem, d["x0"], d["y0"], newton)


@pytest.mark.parametrize("n,m,k", [(16, 8, 0), (32, 16, 4), (24, 0, 5)])
def test_globalized_steps(golden, n, m, k):
    g = golden("globalized")
    key = f"qp_n{n}_m{m}_k{k}"
    problem, _ = make_qp(n, m, k)
    params = orc.OracleParams(newton_type="globalized")
    it = orc.Iterate(problem, params, g[f"{key}/x0"], g[f"{key}/y0"])
    method = orc.newton_method(problem, params, it, float(g[f"{key}/dt"]), float(g[f"{key}/rho"]))
    cur = it
    for j in range(2):
        if f"{key}/s{j}/failed" not in g.files:
            break
        if bool(g[f"{key}/s{j}/failed"]):
            with pytest.raises(orc.LineSearchError):
                method.step(cur)
            break
        res = method.step(cur)
        assert rel_err(res.dx, g[f"{key}/s{j}/dx"]) <= RTOL
        assert rel_err(res.dy, g[f"{key}/s{j}/dy"]) <= RTOL
        assert rel_err(res.iterate.x, g[f"{key}/s{j}/xn"]) <= RTOL
        assert np.array_equal(res.active_set, g[f"{key}/s{j}/active"])
        cur = res.iterate


# ------------------------------------------------------------------ full solves
def _check_solve(g, key, problem, x0, y0, newton="Simplified", tol=RTOL):
    params = orc.OracleParams(newton_type=NEWTON[newton])
    res = orc.Solver(problem, params).solve(x0, y0, record=True)
    assert res.status == int(g[f"{key}/status"])
    horizon = noise_horizon(res.trace)
    accepts = list(g[f"{key}/accepts"])
    assert [t["accept"] for t in res.trace][: horizon + 1] == accepts[: horizon + 1]
    for row, i in enumerate(g[f"{key}/trace_idx"]):
        if i <= horizon and i < len(res.trace):
            assert rel_err(res.trace[i]["x"], g[f"{key}/trace_x"][row]) <= tol, (key, i)
            assert rel_err(res.trace[i]["y"], g[f"{key}/trace_y"][row]) <= tol, (key, i)
    if horizon == len(res.trace):
        assert res.iterations == int(g[f"{key}/iterations"])
        assert res.accepted_steps == int(g[f"{key}/accepted_steps"])
        assert rel_err(res.x, g[f"{key}/x"]) <= tol
        assert rel_err(res.y, g[f"{key}/y"]) <= tol
        assert rel_err(res.d, g[f"{key}/d"]) <= max(tol, 1e-9)
    else:  # past the noise horizon only the converged point is comparable (to opt_tol accuracy)
        assert rel_err(res.x, g[f"{key}/x"]) <= 1e-4
    res.horizon = horizon
    return res


def test_solve_rosenbrock_docs_example(golden):
    """cfg1: docs/solve_rosenbrock.output:5-14 -> 30 iterations, 25 accepted, x=[0.99999959 0.99999917]."""
    g = golden("solves")
    p = orc.ChainedRosenbrock(np.array([1.0]), np.array([100.0]), np.full(2, -np.inf), np.full(2, np.inf))
    res = _check_solve(g, "rosenbrock2d", p, None, None)
    assert res.iterations == 30 and res.accepted_steps == 25
    assert np.array2string(res.x, precision=8) == "[0.99999959 0.99999917]"


@pytest.mark.parametrize("newton", ["Simplified", "Full"])
@pytest.mark.parametrize("S,nx,nu,k", [(6, 3, 2, 0), (16, 4, 3, 1)])
def test_solve_ocp(golden, S, nx, nu, k, newton):
    """cfg4 family (nonlinear equality constraints, bounds on the controls) against the real reference."""
    d = synth.ocp_instance(k, stages=S, nx=nx, nu=nu)
    p = orc.OCP(d["A"], d["B"], d["Q"], d["R"], d["xinit"], d["umax"], d["h"])
    # the 56-iteration trajectory of the larger instance amplifies the last-bit differences of the linear solves
    # (SuperLU's ordering on the sparse vs the dense matrix) to ~1e-10 by iteration 20: compared at 1e-8
    res = _check_solve(golden("ocp"), f"ocp_S{S}_nx{nx}_nu{nu}_k{k}/{newton}", p, d["x0"], d["y0"], newton=newton,
                       tol=RTOL if S == 6 else 1e-8)
    assert res.status == 1


@pytest.mark.parametrize("n,m,k", [(16, 8, 0), (24, 12, 1)])
def test_slack_transform_general_qp(golden, n, m, k):
    """cons_problem.py / transform.py: general constraint bounds through the slack form, against the reference."""
    g = golden("constrained")
    key = f"gqp_n{n}_m{m}_k{k}/Simplified"
    d = synth.general_qp_instance(k, n, m)
    p = orc.GeneralQP(d["H"], d["A"], d["g"], d["b"], d["lb"], d["ub"], d["cons_lb"], d["cons_ub"])
    res = orc.solve_general(p, orc.OracleParams(), d["x0"], d["y0"], record=True)
    assert res.status == int(g[f"{key}/status"]) == 1
    horizon = noise_horizon(res.trace)
    assert [t["accept"] for t in res.trace][: horizon + 1] == list(g[f"{key}/accepts"])[: horizon + 1]
    for row, i in enumerate(g[f"{key}/trace_idx"]):
        if i <= horizon and i < len(res.trace):
            assert rel_err(res.trace[i]["x"], g[f"{key}/trace_x"][row]) <= RTOL  # iterates incl. slacks
    if horizon >= len(res.trace):
        assert res.iterations == int(g[f"{key}/iterations"])
    assert rel_err(res.x, g[f"{key}/x"]) <= 1e-6 and res.x.shape == (n,)
    cons = p.cons(res.x)
    assert (cons >= d["cons_lb"] - 1e-6).all() and (cons <= d["cons_ub"] + 1e-6).all()


def test_slack_transform_hs71_constrained(golden):
    """tests/pygradflow/test_solver.py:133-137 (HS71Constrained): known optimum, and the reference's own trace."""
    g = golden("constrained")
    res = orc.solve_general(orc.HS71Constrained(), orc.OracleParams(), np.array([1.0, 5.0, 5.0, 1.0]), np.zeros(2))
    assert res.status == 1 and res.iterations == int(g["hs71_cons/Simplified/iterations"])
    assert np.allclose(res.x, [1.0, 4.74299964, 3.82114998, 1.37940829], atol=1e-6)  # instances.py:38-40
    assert rel_err(res.x, g["hs71_cons/Simplified/x"]) <= RTOL


@pytest.mark.parametrize("ctl", ["ResiduumRatio", "Exact", "Fixed"])
@pytest.mark.parametrize("prob", ["qp_n16_m8_k0", "qp_n32_m16_k2", "ros_n8_k0"])
def test_other_step_controllers(golden, ctl, prob):
    """residuum_ratio_control.py, exact_control.py, fixed_control.py against the reference's traces."""
    g = golden("controllers")
    key = f"{ctl}/{prob}"
    if prob.startswith("qp"):
        n, m, k = (int(t[1:]) for t in prob.split("_")[1:])
        d = synth.qp_instance(k, n, m)
        p = orc.DenseQP(d["H"], d["A"], d["g"], d["b"], d["lb"], d["ub"])
    else:
        d = synth.rosenbrock_instance(0, 8)
        p = orc.ChainedRosenbrock(d["a"], d["b"], d["lb"], d["ub"])
    name = {"ResiduumRatio": "residuum_ratio", "Exact": "exact", "Fixed": "fixed"}[ctl]
    params = orc.OracleParams(step_control_type=name, iteration_limit=60 if ctl == "Fixed" else None)
    res = orc.Solver(p, params).solve(d["x0"], d["y0"], record=True)
    assert res.status == int(g[f"{key}/status"])
    assert res.iterations == int(g[f"{key}/iterations"]) and res.accepted_steps == int(g[f"{key}/accepted_steps"])
    assert [t["accept"] for t in res.trace] == list(g[f"{key}/accepts"])
    for row, i in enumerate(g[f"{key}/trace_idx"]):
        assert rel_err(res.trace[i]["x"], g[f"{key}/trace_x"][row]) <= 1e-8, (key, i)
    assert rel_err(res.x, g[f"{key}/x"]) <= 1e-8


@pytest.mark.parametrize("kind", ["Asymmetric", "Extended", "Standard"])
@pytest.mark.parametrize("prob", ["Simplified/qp_n16_m8_k0", "Full/qp_n32_m16_k2", "Simplified/qp_n24_m0_k5",
                                  "Full/qp_n24_m0_k5", "Simplified/ros_n8_k0"])
def test_step_solver_formulations(golden, kind, prob):
    """asymmetric_step_solver.py, extended_step_solver.py, standard_step_solver.py against the reference's traces."""
    g = golden("step_solvers")
    key = f"{kind}/{prob}"
    newton, name = prob.split("/")
    if name.startswith("qp"):
        n, m, k = (int(t[1:]) for t in name.split("_")[1:])
        d = synth.qp_instance(k, n, m)
        p = orc.DenseQP(d["H"], d["A"], d["g"], d["b"], d["lb"], d["ub"])
    else:
        d = synth.rosenbrock_instance(0, 8)
        p = orc.ChainedRosenbrock(d["a"], d["b"], d["lb"], d["ub"])
    params = orc.OracleParams(step_solver_type=kind.lower(), newton_type=NEWTON[newton])
    res = orc.Solver(p, params).solve(d["x0"], d["y0"], record=True)
    assert res.status == int(g[f"{key}/status"])
    assert res.iterations == int(g[f"{key}/iterations"]) and res.accepted_steps == int(g[f"{key}/accepted_steps"])
    assert [t["accept"] for t in res.trace] == list(g[f"{key}/accepts"])
    for row, i in enumerate(g[f"{key}/trace_idx"]):
        assert rel_err(res.trace[i]["x"], g[f"{key}/trace_x"][row]) <= 1e-8, (key, i)
    assert rel_err(res.x, g[f"{key}/x"]) <= 1e-8


@pytest.mark.parametrize("newton", ["Simplified", "Full"])
@pytest.mark.parametrize("kind", ["DualEquilibration", "Constant"])
def test_penalty_strategies(golden, kind, newton):
    """penalty.py:36-113 against the reference (iteration_limit = 300).  DualEquilibration drives rho up on these QPs
    until the KKT systems are nearly singular: the reference stops at the iteration limit or at lamb_max, and so does
    the restatement; iterates are compared while they are still rounding-stable (the first 20 iterations)."""
    g = golden("penalty")
    name = {"DualEquilibration": "dual_equilibration", "Constant": "constant"}[kind]
    for (n, m, k) in [(16, 8, 0), (32, 16, 2)]:
        key = f"{kind}/{newton}/qp_n{n}_m{m}_k{k}"
        d = synth.qp_instance(k, n, m)
        p = orc.DenseQP(d["H"], d["A"], d["g"], d["b"], d["lb"], d["ub"])
        params = orc.OracleParams(penalty_update=name, newton_type=NEWTON[newton], iteration_limit=300)
        if bool(g[f"{key}/failed"]):
            with pytest.raises(orc.LambMaxError):
                orc.Solver(p, params).solve(d["x0"], d["y0"])
            continue
        res = orc.Solver(p, params).solve(d["x0"], d["y0"], record=True)
        assert res.status == int(g[f"{key}/status"]) and res.iterations == int(g[f"{key}/iterations"])
        for row, i in enumerate(g[f"{key}/trace_idx"]):
            if i < 20:
                assert rel_err(res.trace[i]["x"], g[f"{key}/trace_x"][row]) <= 1e-8, (key, i)
        if kind == "Constant":
            assert [t["accept"] for t in res.trace] == list(g[f"{key}/accepts"])
            assert rel_err(res.x, g[f"{key}/x"]) <= 1e-8


SCALING = {"GradJac": "grad_jac", "KKT": "kkt", "Nominal": "nominal", "Custom": "custom"}


def _scaling_case(g, name, kind):
    key = f"{name}/{kind}"
    if name.startswith("gqp"):
        n, m, k = (int(t[1:]) for t in name.split("_")[1:])
        d = synth.general_qp_instance(k, n, m)
        p = orc.GeneralQP(d["H"], d["A"], d["g"], d["b"], d["lb"], d["ub"], d["cons_lb"], d["cons_ub"])
        x0, y0 = d["x0"], d["y0"]
    else:
        p, x0, y0 = orc.HS71Constrained(), np.array([1.0, 5.0, 5.0, 1.0]), np.zeros(2)
    sc = None
    if kind == "Custom":
        sc = orc.Scaling(g[f"{key}/var_weights"], g[f"{key}/cons_weights"], int(g[f"{key}/obj_weight"]))
    params = orc.OracleParams(scaling_type=SCALING[kind], scaling=sc, scaling_primal=g[f"{key}/scaling_primal"],
                              scaling_dual=g[f"{key}/scaling_dual"])
    return key, p, params, x0, y0


@pytest.mark.parametrize("kind", list(SCALING))
@pytest.mark.parametrize("name", ["gqp_n16_m8_k0", "gqp_n24_m12_k1", "hs71_cons"])
def test_scaling_weights_and_solve(golden, name, kind):
    """scale.py (create_scaling, ScaledProblem, Transformation.transform_sol / restore_sol) against the reference:
    identical integer weights; same status and solution.  On the well-conditioned HS71 runs the whole trace and
    the iteration counts are identical; the scaled random QPs start with a dozen rejected steps on nearly singular
    systems (the oracle with LAPACK instead of SuperLU already takes a different path there), so only what survives
    that is compared."""
    g = golden("scaling")
    key, p, params, x0, y0 = _scaling_case(g, name, kind)
    sc = orc.create_scaling(p, params, params.scaling_primal, params.scaling_dual)
    assert np.array_equal(sc.var_weights, g[f"{key}/var_weights"])
    assert np.array_equal(sc.cons_weights, g[f"{key}/cons_weights"])
    assert sc.obj_weight == int(g[f"{key}/obj_weight"])
    res = orc.solve_general(p, params, x0, y0, record=True)
    assert res.status == int(g[f"{key}/status"]) == 1
    assert rel_err(res.x, g[f"{key}/x"]) <= 2e-6
    assert rel_err(res.y, g[f"{key}/y"]) <= 1e-4
    if name == "hs71_cons" and kind != "Nominal":
        assert res.iterations == int(g[f"{key}/iterations"])
        assert [t["accept"] for t in res.trace] == list(g[f"{key}/accepts"])
        for row, i in enumerate(g[f"{key}/trace_idx"]):
            assert rel_err(res.trace[i]["x"], g[f"{key}/trace_x"][row]) <= 1e-8, (key, i)
        assert rel_err(res.d, g[f"{key}/d"]) <= 1e-6


def test_scaled_problem_callbacks_exact():
    """ScaledProblem (scale.py:153-231) is exact in floating point (ldexp): the scaled callbacks equal the unscaled
    ones times the corresponding powers of two, bit for bit; zero weights are the identity (test_scale.py:22-48)."""
    p = orc.HS71Constrained()
    rng = np.random.default_rng(0)
    sc = orc.Scaling(np.array([1, 2, 3, -4]), np.array([3, 2]), 4)      # test_scale.py:51-60 (first four weights)
    sp_ = orc.ScaledProblem(p, sc)
    x, y = rng.uniform(1.0, 5.0, 4), rng.standard_normal(2)
    xs, ys = sc.scale_primal(x), sc.scale_dual(y)
    assert sp_.obj(xs) == p.obj(x) * 2.0 ** 4                           # test_scale.py:63-80
    assert np.array_equal(sp_.obj_grad(xs), p.obj_grad(x) * 2.0 ** (4 - sc.var_weights))
    assert np.array_equal(sp_.cons(xs), p.cons(x) * 2.0 ** sc.cons_weights)
    J = np.asarray(p.cons_jac(x).todense() if hasattr(p.cons_jac(x), "todense") else p.cons_jac(x))
    assert np.array_equal(sp_.cons_jac(xs), J * 2.0 ** (sc.cons_weights[:, None] - sc.var_weights[None, :]))
    z = orc.ScaledProblem(p, orc.Scaling(np.zeros(4, dtype=int), np.zeros(2, dtype=int)))
    assert z.obj(x) == p.obj(x) and np.array_equal(z.obj_grad(x), p.obj_grad(x))
    assert np.array_equal(z.lag_hess(x, y), np.asarray(p.lag_hess(x, y)))
    assert np.array_equal(sc.unscale_primal(xs), x) and np.array_equal(sc.unscale_dual(ys), y)


@pytest.mark.parametrize("newton", ["Simplified", "Full"])
@pytest.mark.parametrize("kind", ["Smallest", "Explicit"])
def test_tau_active_set_types(golden, kind, newton):
    """newton_control.py:40-88 + implicit_func.py:237-244 against the reference."""
    g = golden("active_set_types")
    okw = {"Smallest": dict(active_set_type="smallest"), "Explicit": dict(active_set_type="explicit", active_set_tau=0.3)}[kind]
    for (n, m, k) in [(16, 8, 0), (32, 16, 2)]:
        d = synth.qp_instance(k, n, m)
        p = orc.DenseQP(d["H"], d["A"], d["g"], d["b"], d["lb"], d["ub"])
        res = orc.Solver(p, orc.OracleParams(newton_type=NEWTON[newton], **okw)).solve(d["x0"], d["y0"], record=True)
        key = f"{kind}/{newton}/qp_n{n}_m{m}_k{k}"
        assert res.status == int(g[f"{key}/status"]) and res.iterations == int(g[f"{key}/iterations"])
        assert [t["accept"] for t in res.trace] == list(g[f"{key}/accepts"])
        assert rel_err(res.x, g[f"{key}/x"]) <= RTOL


def test_solve_tame(golden):
    res = _check_solve(golden("solves"), "tame", orc.Tame(), np.zeros(2), np.zeros(1))
    assert np.allclose(res.x, [0.5, 0.5], atol=1e-6)  # tests/pygradflow/instances.py:57-68


@pytest.mark.parametrize("newton", list(NEWTON))
def test_solve_hs71(golden, newton):
    res = _check_solve(
        golden("solves"), f"hs71/{newton}", orc.HS71(), np.array([1.0, 5.0, 5.0, 1.0, 0.0]), np.zeros(2), newton
    )
    # tests/pygradflow/instances.py:38-40
    assert np.allclose(res.x, [1.0, 4.74299964, 3.82114998, 1.37940829, 0.0], atol=1e-6)
    assert np.allclose(res.y, [-0.55229366, 0.16146857], atol=1e-6)


@pytest.mark.parametrize(
    "n,m,k,newton",
    [(16, 8, 0, "Simplified"), (16, 8, 0, "Full"), (16, 8, 0, "ActiveSet"), (16, 8, 1, "Simplified"),
     (16, 8, 1, "Full"), (32, 16, 2, "Simplified"), (64, 32, 3, "Simplified"), (64, 32, 4, "Simplified"),
     (48, 0, 5, "Simplified")],
)
def test_solve_qp(golden, n, m, k, newton):
    problem, d = make_qp(n, m, k)
    _check_solve(golden("solves"), f"qp_n{n}_m{m}_k{k}/{newton}", problem, d["x0"], d["y0"], newton)


@pytest.mark.parametrize("n,k", [(8, 0), (8, 1), (16, 2), (64, 1)])
def test_solve_chained_rosenbrock(golden, n, k):
    problem, d = make_ros(n, k)
    _check_solve(golden("solves"), f"ros_n{n}_k{k}/Simplified", problem, d["x0"], d["y0"], tol=1e-9)


def test_solve_qp_full_size(golden):
    """One cfg3-size instance (n=512, m=256)."""
    problem, d = make_qp(512, 256, 0)
    _check_solve(golden("qp512"), "qp_n512_m256_k0/Simplified", problem, d["x0"], d["y0"])


# ------------------------------------------------------------------ reference unit-test restatements
def test_identity_limit_steps():
    """tests/pygradflow/test_newton.py:142-214: dt -> 0 gives F' ~ I and x+ ~ clip(x)."""
    for newton in ("simplified", "full", "active_set", "globalized"):
        p = orc.ChainedRosenbrock(np.array([1.0]), np.array([100.0]), np.full(2, -np.inf), np.full(2, np.inf))
        params = orc.OracleParams(newton_type=newton)
        it = orc.Iterate(p, params, np.zeros(2), np.zeros(0))
        f = orc.ImplicitFunc(p, it, 1e-10)
        assert np.allclose(f.deriv_at(it, 1.0), np.eye(2))
        nxt = orc.newton_method(p, params, it, 1e-10, 1.0).step(it).iterate
        assert np.allclose(nxt.x, it.x)
        # everything active (:176-214)
        p2 = orc.ChainedRosenbrock(np.array([1.0]), np.array([100.0]), np.ones(2), np.array([1.0, np.inf]))
        it2 = orc.Iterate(p2, params, np.zeros(2), np.zeros(0))
        nxt2 = orc.newton_method(p2, params, it2, 1e-12, 1.0).step(it2).iterate
        assert np.allclose(nxt2.x, np.clip(it2.x, p2.var_lb, p2.var_ub))


def test_one_step_convergence_tame():
    """tests/pygradflow/test_solver.py:191-215."""
    for newton in ("simplified", "full", "active_set", "globalized"):
        p = orc.Tame()
        params = orc.OracleParams(newton_type=newton)
        it = orc.Iterate(p, params, np.zeros(2), np.zeros(1))
        nxt = orc.newton_method(p, params, it, 10.0, 1.0).step(it).iterate
        assert np.allclose(orc.ImplicitFunc(p, it, 10.0).value_at(nxt, 1.0), 0.0)


def test_func_zero_at_origin():
    """tests/pygradflow/test_func.py:10-26: F(x^, y^) -> 0 as dt -> 0."""
    p = orc.HS71()
    it = orc.Iterate(p, orc.OracleParams(), np.array([1.0, 5.0, 5.0, 1.0, 0.0]), np.zeros(2))
    assert np.allclose(orc.ImplicitFunc(p, it, 1e-12).value_at(it, 1.0), 0.0, atol=1e-8)


@pytest.mark.parametrize("name", ["indef", "posdef", "negdef", "kkt12", "kkt48", "kkt96", "kkt200"])
def test_condition_estimator(golden, name):
    """step/cond_estimate.py:13-114 (Dixon) on the reference's LUSolver: same estimate, to the last bits."""
    g, ls = golden("rcond"), golden("linear_solver")
    mat = ls[f"{name}/mat"] if not name.startswith("kkt") else synth.kkt_instance(int(name[3:]))[0]
    est = orc.ConditionEstimator(mat, orc.OracleLUSolver(mat))
    assert est.required_its() == int(g[f"{name}/its"])
    assert abs(est.estimate_rcond() - float(g[f"{name}/rcond"])) <= 1e-13 * float(g[f"{name}/rcond"])


def check_iterative_trace(res, g, key, prefix=10):
    """The iterative solvers stop at rtol = 1e-5, so a Newton step is only defined up to the iteration count of the
    Krylov method: two implementations whose matrix-vector products round differently (sparse CSC here and in the
    reference, but with different index orders; a block tree on the GPU) agree to rounding while every solve stops
    after the same number of products and differ at the 1e-5 level afterwards.  Checked: the same status, the first
    `prefix` outer iterations identical (accept sequence, iterates within 1e-8), the same optimum (to opt_tol)."""
    assert res.status == int(g[f"{key}/status"])
    acc = list(g[f"{key}/accepts"])
    k = min(prefix, len(acc), len(res.trace))
    assert [t["accept"] for t in res.trace][:k] == acc[:k]
    for row, i in enumerate(g[f"{key}/trace_idx"]):
        if i < k and res.trace[i]["accept"]:  # the candidate of a step that died with a LinearSolverError is never used
            assert rel_err(res.trace[i]["x"], g[f"{key}/trace_x"][row]) <= 1e-8, (key, i)
    assert abs(res.iterations - int(g[f"{key}/iterations"])) <= max(3, int(g[f"{key}/iterations"]) // 5)
    assert rel_err(res.x, g[f"{key}/x"]) <= 1e-5


ITER_MATS = ["indef", "posdef", "negdef", "kkt12", "kkt48", "kkt96", "unsym10", "unsym40"]


@pytest.mark.parametrize("name", ITER_MATS)
def test_iterative_linear_solvers(golden, name):
    """gmres_solver.py / minres_solver.py: the oracle's wrappers against the reference's own (same scipy routine)."""
    g = golden("iterative")
    mat, rhs, x0 = g[f"{name}/mat"], g[f"{name}/rhs"], g[f"{name}/x0"]
    sym = bool(np.array_equal(mat, mat.T))
    for kind in ("gmres", "minres"):
        if kind == "minres" and not sym:
            continue
        s = orc.make_linear_solver(mat, kind, symmetric=sym)
        cases = [("sol", {}), ("sol_x0", dict(initial_sol=lambda: x0.copy()))]
        if kind == "gmres":
            cases.append(("sol_trans", dict(trans=True)))
        for key, kw in cases:
            if bool(g[f"{name}/{kind}/{key}_failed"]):
                with pytest.raises(orc.LinearSolverError):
                    s.solve(rhs, **kw)
            else:
                assert np.array_equal(s.solve(rhs, **kw), g[f"{name}/{kind}/{key}"]), (name, kind, key)


@pytest.mark.parametrize("newton", ["Simplified", "Full"])
@pytest.mark.parametrize("lin,form", [("GMRES", "Symmetric"), ("GMRES", "Asymmetric"), ("GMRES", "Extended"),
                                      ("MINRES", "Symmetric")])
def test_iterative_solves(golden, lin, form, newton):
    """Solver.solve with Params.linear_solver_type = GMRES / MINRES against the reference's traces."""
    g = golden("iterative")
    for (n, m, k) in [(16, 8, 0), (32, 16, 2), (24, 0, 5)]:
        key = f"{lin}/{form}/{newton}/qp_n{n}_m{m}_k{k}"
        d = synth.qp_instance(k, n, m)
        p = orc.DenseQP(d["H"], d["A"], d["g"], d["b"], d["lb"], d["ub"])
        params = orc.OracleParams(linear_solver=lin.lower(), step_solver_type=form.lower(), newton_type=NEWTON[newton],
                                  iteration_limit=400)
        if bool(g[f"{key}/failed"]):
            with pytest.raises(orc.LambMaxError):
                orc.Solver(p, params).solve(d["x0"], d["y0"])
            continue
        res = orc.Solver(p, params).solve(d["x0"], d["y0"], record=True)
        check_iterative_trace(res, g, key)


PENALTY_REJECT_CASES = ["qp_n16_m8_k0", "qp_n32_m16_k2", "hs71", "tame"] + [f"qpr_n12_m5_k{k}" for k in
                                                                              (0, 2, 3, 10, 18, 32, 35)]
PENALTY_KIND = {"ParetoDecrease": "pareto_decrease", "ObjectiveFilter": "objective_filter",
                "LagrangianFilter": "lagrangian_filter"}


def penalty_reject_problem(name):
    """(oracle problem, data dict or None, x0, y0) of a case of tests/golden/penalty_reject.npz."""
    if name == "hs71":
        return orc.HS71(), None, np.array([1.0, 5.0, 5.0, 1.0, 0.0]), np.zeros(2)
    if name == "tame":
        return orc.Tame(), None, np.zeros(2), np.zeros(1)
    fam, n, m, k = name.split("_")
    n, m, k = int(n[1:]), int(m[1:]), int(k[1:])
    d = synth.qp_instance(k, n, m)
    x0, y0 = d["x0"], d["y0"]
    if fam == "qpr":
        rng = np.random.default_rng(500 + k)
        x0, y0 = rng.uniform(-1, 1, n), rng.normal(size=m) * 2
    return orc.DenseQP(d["H"], d["A"], d["g"], d["b"], d["lb"], d["ub"]), d, x0, y0


@pytest.mark.parametrize("newton", ["Simplified", "Full"])
@pytest.mark.parametrize("kind", list(PENALTY_KIND))
def test_penalty_strategies_that_see_the_candidate(golden, kind, newton):
    """penalty.py:115-255 (ParetoDecrease, ObjectiveFilter, LagrangianFilter) + solver.py:357-378 against the real
    reference: status, iteration / accepted counts, the accept sequence (incl. filter rejections) and the rho the
    solver used in every iteration."""
    g = golden("penalty_reject")
    for name in PENALTY_REJECT_CASES:
        key = f"{kind}/{newton}/{name}"
        assert not bool(g[f"{key}/failed"])
        p, _, x0, y0 = penalty_reject_problem(name)
        params = orc.OracleParams(penalty_update=PENALTY_KIND[kind], newton_type=NEWTON[newton], iteration_limit=300)
        res = orc.Solver(p, params).solve(x0, y0, record=True)
        assert res.status == int(g[f"{key}/status"]), key
        assert res.iterations == int(g[f"{key}/iterations"]) and res.accepted_steps == int(g[f"{key}/accepted_steps"]), key
        got = [t["accept"] and t.get("penalty_accept", True) for t in res.trace]
        # the reference's callback sees the controller's verdict (before the penalty strategy may veto it)
        assert [t["accept"] for t in res.trace] == list(g[f"{key}/accepts"]), key
        assert sum(got) == res.accepted_steps
        rhos = np.array([t["rho"] for t in res.trace])
        assert np.allclose(rhos, g[f"{key}/rhos"], rtol=1e-9, atol=0.0), key
        assert rel_err(res.x, g[f"{key}/x"]) <= 1e-8, key
