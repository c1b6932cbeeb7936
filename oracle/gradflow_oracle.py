"""CPU oracle for the pygradflow inner Newton/KKT path.  TEST INFRASTRUCTURE ONLY.

This module is a dense-NumPy restatement of the reference algorithm
(chrhansk/pygradflow v0.5.24) for the hot path named in BASELINE.json.  It is
*not* product code: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.
The product package ``pygradflow_b200`` never imports anything from here.

Parity status: PINNED.  ``tests/golden/make_golden.py`` runs the real reference
(imported from ``/root/reference`` in the build container) on the synthetic
problem families and on the reference's own test fixtures and stores the traces
under ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks this
restatement against those traces (iterates <= 1e-10 rel, identical active sets,
accept sequences, iteration counts and status).

The factorisation arithmetic in the reference is SuperLU inside SciPy
(``scipy.sparse.linalg.splu``, scipy unpinned ">=1.14", 1.18.1 installed; call
sites ``pygradflow/linear_solver/lu_solver.py:14,21``).  The oracle calls the
same routine on the same CSC matrix (``linear_solver="splu"``, default) so it is
a faithful CPU baseline; ``linear_solver="lapack"`` uses dense ``getrf`` instead.

Every function cites the reference ``file:line`` it follows (paths relative to
``/root/reference/``).
"""

from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Any, Callable, List, Optional

import numpy as np
import scipy.linalg
import scipy.sparse
import scipy.sparse.linalg

# --------------------------------------------------------------------------
# status words (pygradflow/status.py:4-31) -- values are the enum's auto() ints
# --------------------------------------------------------------------------
STATUS_OPTIMAL = 1
STATUS_ITERATION_LIMIT = 2
STATUS_TIME_LIMIT = 3
STATUS_UNBOUNDED = 4
STATUS_LOCALLY_INFEASIBLE = 5
# not a reference status: Solver.solve raises when lamb >= lamb_max
# (pygradflow/solver.py:323-326); the batched driver needs a word for it.
STATUS_LAMB_MAX = 6
STATUS_LINE_SEARCH_FAILED = 7

STATUS_NAMES = {
    0: "running",
    1: "optimal",
    2: "iteration_limit",
    3: "time_limit",
    4: "unbounded",
    5: "infeasible",
    6: "lamb_max_exceeded",
    7: "line_search_failed",
}


class LinearSolverError(Exception):
    """pygradflow/linear_solver/linear_solver.py:8-15."""


class StepSolverError(Exception):
    """pygradflow/step/step_solver_error.py:1-7."""


class LambMaxError(Exception):
    """Raised where pygradflow/solver.py:323-326 raises a bare Exception."""


class LineSearchError(Exception):
    """Raised where pygradflow/newton.py:294 raises a bare Exception."""


@dataclass
class OracleParams:
    """The subset of pygradflow/params.py:197-265 the hot path reads (same names, same defaults)."""

    rho: float = 1e-8
    theta_max: float = 0.9
    theta_ref: float = 0.5
    lamb_init: float = 1.0
    lamb_min: float = 1e-12
    lamb_max: float = 1e12
    lamb_inc: float = 2.0
    lamb_red: float = 0.5
    K_P: float = 0.2
    K_I: float = 0.005
    opt_tol: float = 1e-6
    active_tol: float = 1e-8
    local_infeas_tol: float = 1e-8
    newton_type: str = "simplified"  # simplified | full | active_set | globalized
    newton_tol: float = 1e-8
    step_control_type: str = "distance_ratio"  # distance_ratio | residuum_ratio | exact | fixed
    active_set_type: str = "standard"  # standard | explicit | smallest | largest  (params.py:14-18,225-227)
    active_set_tau: Optional[float] = None
    # dual_norm | constant | dual_equilibration | pareto_decrease | objective_filter | lagrangian_filter
    penalty_update: str = "dual_norm"
    iteration_limit: Optional[int] = None
    obj_lower_limit: float = -1e10
    inertia_correction: bool = False
    linear_solver: str = "splu"  # splu | lapack
    # plugin hook, same contract as pygradflow/step/solver/__init__.py:18-19
    step_solver: Optional[Callable] = None
    step_solver_type: str = "symmetric"  # symmetric | asymmetric | extended | standard  (params.py:50-70,235)
    scaling_type: str = "none"  # none | grad_jac | kkt | nominal | custom  (params.py:166-195,245-250)
    scaling: Optional[Any] = None
    scaling_primal: Optional[np.ndarray] = None
    scaling_dual: Optional[np.ndarray] = None
    # keep J / H / K as scipy.sparse matrices like the reference does (csr J, csc K; symmetric_step_solver.py:27-94)
    # instead of dense arrays -- same arithmetic, needed for cfg4-sized banded problems (Symmetric formulation only)
    sparse: bool = False
    penalty_filter_capacity: Optional[int] = None  # unused by the oracle (its filter is an unbounded list)
    dtype = np.float64


def _dense(a) -> np.ndarray:
    if scipy.sparse.issparse(a):
        return np.asarray(a.toarray(), dtype=np.float64)
    return np.asarray(a, dtype=np.float64)


# --------------------------------------------------------------------------
# Problem base (pygradflow/problem.py:32-192) and families
# --------------------------------------------------------------------------
class OracleProblem:
    """Equality-constrained (c(x)=0) + bound-constrained problem, reference callback names."""

    def __init__(self, var_lb, var_ub, num_cons=0):
        self.var_lb = np.array(var_lb, dtype=np.float64)
        self.var_ub = np.array(var_ub, dtype=np.float64)
        assert self.var_lb.shape == self.var_ub.shape and self.var_lb.ndim == 1
        assert (self.var_lb <= self.var_ub).all()
        self.num_cons = int(num_cons)
        self.cons_lb = np.zeros(self.num_cons)
        self.cons_ub = np.zeros(self.num_cons)

    @property
    def num_vars(self):
        return self.var_lb.shape[0]

    def obj(self, x):
        raise NotImplementedError

    def obj_grad(self, x):
        raise NotImplementedError

    def cons(self, x):
        return np.zeros(0)

    def cons_jac(self, x):
        return np.zeros((0, self.num_vars))

    def lag_hess(self, x, y):
        raise NotImplementedError


class GeneralQP(OracleProblem):
    """QP with general constraints cl <= Ax + b <= cu (rows with cl == cu are equalities): the form the
    reference's ConstrainedProblem (cons_problem.py:8-173) turns into equalities + slack bounds."""

    def __init__(self, H, A, g, b, lb, ub, cons_lb, cons_ub):
        super().__init__(lb, ub, num_cons=np.asarray(A).shape[0])
        self.H, self.A = _dense(H), _dense(A)
        self.g, self.b = np.asarray(g, dtype=np.float64), np.asarray(b, dtype=np.float64)
        self.cons_lb = np.asarray(cons_lb, dtype=np.float64)
        self.cons_ub = np.asarray(cons_ub, dtype=np.float64)

    def obj(self, x):
        return float(0.5 * x @ (self.H @ x) + self.g @ x)

    def obj_grad(self, x):
        return self.H @ x + self.g

    def cons(self, x):
        return self.A @ x + self.b

    def cons_jac(self, x):
        return self.A

    def lag_hess(self, x, y):
        return self.H


class HS71Constrained(OracleProblem):
    """tests/pygradflow/hs71_cons.py: HS71 with an inequality and an equality constraint (no explicit slack)."""

    def __init__(self):
        super().__init__(np.ones(4), np.full(4, 5.0), num_cons=2)
        self.cons_lb = np.array([25.0, 40.0])
        self.cons_ub = np.array([np.inf, 40.0])

    def obj(self, x):
        return x[0] * x[3] * (x[0] + x[1] + x[2]) + x[2]

    def obj_grad(self, x):
        return np.array([(x[0] + x[1] + x[2]) * x[3] + x[0] * x[3], x[0] * x[3], x[0] * x[3] + 1,
                         (x[0] + x[1] + x[2]) * x[0]])

    def cons(self, x):
        return np.array([np.prod(x), np.dot(x, x)])

    def cons_jac(self, x):
        return np.array([[x[1] * x[2] * x[3], x[0] * x[2] * x[3], x[0] * x[1] * x[3], x[0] * x[1] * x[2]],
                         [2 * x[0], 2 * x[1], 2 * x[2], 2 * x[3]]])

    def lag_hess(self, x, lag):
        l1, l2 = lag
        oh = np.array([[2 * x[3], x[3], x[3], 2 * x[0] + x[1] + x[2]], [x[3], 0, 0, x[0]], [x[3], 0, 0, x[0]],
                       [2 * x[0] + x[1] + x[2], x[0], x[0], 0]])
        h1 = np.array([[0, x[2] * x[3], x[1] * x[3], x[1] * x[2]], [x[2] * x[3], 0, x[0] * x[3], x[0] * x[2]],
                       [x[1] * x[3], x[0] * x[3], 0, x[0] * x[1]], [x[1] * x[2], x[0] * x[2], x[0] * x[1], 0]])
        return oh + l1 * h1 + l2 * 2.0 * np.eye(4)


class ConstrainedProblem(OracleProblem):
    """Slack reformulation of cl <= c(x) <= cu into equalities + bounds (cons_problem.py:8-173): a slack per
    inequality row, c_i(x) - s_i = 0 with cl_i <= s_i <= cu_i; equality rows are shifted by -cl_i."""

    def __init__(self, problem):
        self.problem = problem
        cl, cu = problem.cons_lb, problem.cons_ub
        eq = cl == cu                                                  # :38-46
        self.slack_positions = np.flatnonzero(~eq)
        offs = np.where(eq & (cl != 0.0), -cl, 0.0)
        self.cons_offsets = offs if (eq & (cl != 0.0)).any() else None
        sp = self.slack_positions
        super().__init__(np.concatenate([problem.var_lb, cl[sp]]), np.concatenate([problem.var_ub, cu[sp]]),
                         num_cons=problem.num_cons)                    # :14-29

    def _orig(self, x):
        return x[: self.problem.num_vars]

    def obj(self, x):
        return self.problem.obj(self._orig(x))

    def obj_grad(self, x):
        return np.concatenate([self.problem.obj_grad(self._orig(x)), np.zeros(len(self.slack_positions))])

    def cons(self, x):                                                 # :77-94
        c = np.array(self.problem.cons(self._orig(x)), dtype=np.float64, copy=True)
        if self.cons_offsets is not None:
            c += self.cons_offsets
        c[self.slack_positions] -= x[self.problem.num_vars:]
        return c

    def cons_jac(self, x):                                             # :96-113
        J = _dense(self.problem.cons_jac(self._orig(x)))
        E = np.zeros((self.num_cons, len(self.slack_positions)))
        E[self.slack_positions, np.arange(len(self.slack_positions))] = -1.0
        return np.hstack([J, E])

    def lag_hess(self, x, y):                                          # :115-128
        H = _dense(self.problem.lag_hess(self._orig(x), y))
        n, ns = H.shape[0], len(self.slack_positions)
        out = np.zeros((n + ns, n + ns))
        out[:n, :n] = H
        return out

    def transform_sol(self, x, y):                                     # :130-157
        c = self.problem.cons(x)
        sp = self.slack_positions
        return np.concatenate([x, np.clip(c[sp], self.problem.cons_lb[sp], self.problem.cons_ub[sp])]), y

    def restore_sol(self, x, y, d):                                    # :159-173
        return self._orig(x), y, self._orig(d)


# --------------------------------------------------------------------------
# Power-of-two scaling (pygradflow/scale.py)
# --------------------------------------------------------------------------
def scale_symmetric(A):
    """scale.py:12-45: iterative power-of-two equilibration of a symmetric matrix; returns integer exponents D.
    The reference accumulates the column sums in an INTEGER array (`R = np.zeros(n, dtype=int); R[c] += |a|`), i.e.
    every partial sum is truncated: restated literally."""
    A = np.asarray(A, dtype=np.float64)
    n = A.shape[0]
    rows, cols = np.nonzero(A)                      # COO order of tocoo() on a dense-built matrix: row-major
    data = np.abs(A[rows, cols])
    D = np.zeros(n, dtype=int)
    for _ in range(100):
        R = np.zeros(n, dtype=int)
        for k in range(len(data)):
            R[cols[k]] += data[k]
        R = R.astype(np.float64)
        R[R < 1e-10] = 1.0
        R = np.sqrt(R)
        Rsca = 1 - np.frexp(R)[1]
        if (Rsca == 0).all():
            break
        data = np.ldexp(data, Rsca[rows] + Rsca[cols])
        D += Rsca
    else:
        raise Exception("Equilibration failed to converge")
    return D


class Scaling:
    """scale.py:48-150: integer exponents; x_scaled = ldexp(x, var_weights), c_scaled = ldexp(c, cons_weights)."""

    def __init__(self, var_weights, cons_weights, obj_weight=0):
        self.var_weights = np.asarray(var_weights).astype(int)
        self.cons_weights = np.asarray(cons_weights).astype(int)
        self.obj_weight = int(obj_weight)

    @staticmethod
    def weights_from_nominal_values(values):                           # :75-77
        return 1 - np.frexp(values)[1]

    @staticmethod
    def from_nominal_values(var_values, cons_values, obj_value=1.0):   # :67-73
        w = Scaling.weights_from_nominal_values
        return Scaling(w(var_values), w(cons_values), w(obj_value))

    @staticmethod
    def from_grad_jac(obj_grad, cons_jac):                             # :79-106
        var_weights = -Scaling.weights_from_nominal_values(np.abs(obj_grad))
        if cons_jac is None:
            return Scaling(var_weights, np.zeros(0, dtype=int))
        J = np.abs(_dense(cons_jac))
        pres = np.ldexp(J, np.broadcast_to(-var_weights, J.shape))
        max_values = np.zeros(J.shape[0], dtype=int)                   # integer array in the reference: truncation
        for i in range(J.shape[0]):
            for j in range(J.shape[1]):
                if J[i, j] != 0.0:
                    max_values[i] = max(max_values[i], pres[i, j])
        return Scaling(var_weights, Scaling.weights_from_nominal_values(max_values))

    @staticmethod
    def from_equilibrated_kkt(lag_hess, cons_jac):                     # :108-119
        H, J = _dense(lag_hess), _dense(cons_jac)
        m, n = J.shape
        K = np.zeros((n + m, n + m))
        K[:n, :n], K[:n, n:], K[n:, :n] = H, J.T, J
        w = scale_symmetric(K)
        return Scaling(-w[:n], w[n:])

    def scale_primal(self, x):
        return np.ldexp(x, self.var_weights)

    def unscale_primal(self, x):
        return np.ldexp(x, -self.var_weights)

    def scale_dual(self, y):
        return np.ldexp(y, -(self.cons_weights - self.obj_weight))

    def unscale_dual(self, y):
        return np.ldexp(y, self.cons_weights - self.obj_weight)

    def unscale_bounds_dual(self, d):
        return np.ldexp(d, self.var_weights - self.obj_weight)


class ScaledProblem(OracleProblem):
    """scale.py:153-231."""

    def __init__(self, problem, scaling):
        self.problem, self.scaling = problem, scaling
        vw, cw = scaling.var_weights, scaling.cons_weights
        super().__init__(np.ldexp(problem.var_lb, vw), np.ldexp(problem.var_ub, vw), num_cons=problem.num_cons)
        self.cons_lb = np.ldexp(problem.cons_lb, cw)
        self.cons_ub = np.ldexp(problem.cons_ub, cw)

    def _orig_x(self, x):
        return np.ldexp(x, -self.scaling.var_weights)

    def obj(self, x):
        return np.ldexp(self.problem.obj(self._orig_x(x)), self.scaling.obj_weight)

    def obj_grad(self, x):
        g = np.ldexp(self.problem.obj_grad(self._orig_x(x)), -self.scaling.var_weights)
        return np.ldexp(g, self.scaling.obj_weight)

    def cons(self, x):
        return np.ldexp(self.problem.cons(self._orig_x(x)), self.scaling.cons_weights)

    def cons_jac(self, x):
        J = _dense(self.problem.cons_jac(self._orig_x(x)))
        vw, cw = self.scaling.var_weights, self.scaling.cons_weights
        return np.ldexp(J, cw[:, None] - vw[None, :])

    def lag_hess(self, x, y):
        s = self.scaling
        H = _dense(self.problem.lag_hess(self._orig_x(x), np.ldexp(y, s.cons_weights - s.obj_weight)))
        return np.ldexp(H, s.obj_weight - s.var_weights[:, None] - s.var_weights[None, :])


def create_scaling(problem, params, scaling_primal=None, scaling_dual=None):
    """scale.py:234-280.  params.scaling_type: none | custom | nominal | grad_jac | kkt."""
    st = params.scaling_type
    if params.scaling is not None:
        assert st == "custom"
        return params.scaling
    if st == "none":
        return None
    if st == "custom":
        raise ValueError("Custom scaling requires explicit scaling")
    if scaling_primal is None:
        raise ValueError("Primal point required for scaling computation")
    if st == "nominal":
        cons_val = problem.cons(scaling_primal) if problem.num_cons > 0 else np.zeros(0)
        return Scaling.from_nominal_values(scaling_primal, cons_val)
    cons_jac = problem.cons_jac(scaling_primal) if problem.num_cons > 0 else np.zeros((0, problem.num_vars))
    if st == "grad_jac":
        return Scaling.from_grad_jac(problem.obj_grad(scaling_primal), cons_jac)
    if st == "kkt":
        if scaling_dual is None:
            raise ValueError("Dual point required for KKT scaling computation")
        return Scaling.from_equilibrated_kkt(problem.lag_hess(scaling_primal, scaling_dual), cons_jac)
    raise ValueError(f"Unknown scaling type {st}")


def solve_general(problem, params=None, x0=None, y0=None, record=False):
    """Solver.solve for a problem with general constraint bounds: Transformation (transform.py:13-104) = optional
    power-of-two scaling (scale.py), slack transform, create_transformed_iterate (:29-54), the solve, restore_sol
    (:90-104)."""
    params = params if params is not None else OracleParams()
    n, m = problem.num_vars, problem.num_cons
    x = np.clip(np.zeros(n), problem.var_lb, problem.var_ub) if x0 is None else np.broadcast_to(x0, (n,)).astype(float)
    y = np.zeros(m) if y0 is None else np.broadcast_to(y0, (m,)).astype(float)
    scaling = create_scaling(problem, params, params.scaling_primal, params.scaling_dual)
    scaled = problem if scaling is None else ScaledProblem(problem, scaling)
    if scaling is not None:
        x, y = scaling.scale_primal(x), scaling.scale_dual(y)
    cp = ConstrainedProblem(scaled)
    xt, yt = cp.transform_sol(x, y)
    res = Solver(cp, params).solve(xt, yt, record=record)
    res.x_slack = res.x
    res.x, res.y, res.d = cp.restore_sol(res.x, res.y, res.d)
    if scaling is not None:
        res.x, res.y, res.d = scaling.unscale_primal(res.x), scaling.unscale_dual(res.y), scaling.unscale_bounds_dual(res.d)
    res.scaling = scaling
    return res


class DenseQP(OracleProblem):
    """f = 1/2 x'Hx + g'x, c = Ax + b  (same formulas as tests/pygradflow/qp.py:4-30)."""

    def __init__(self, H, A, g, b, lb, ub):
        super().__init__(lb, ub, num_cons=A.shape[0])
        self.H = np.asarray(H, dtype=np.float64)
        self.A = np.asarray(A, dtype=np.float64)
        self.g = np.asarray(g, dtype=np.float64)
        self.b = np.asarray(b, dtype=np.float64)

    def obj(self, x):
        return 0.5 * x @ (self.H @ x) + self.g @ x

    def obj_grad(self, x):
        return self.H @ x + self.g

    def cons(self, x):
        return self.A @ x + self.b

    def cons_jac(self, x):
        return self.A

    def lag_hess(self, x, y):
        return self.H


class ChainedRosenbrock(OracleProblem):
    """f = sum_i b_i (x_{i+1} - x_i^2)^2 + (a_i - x_i)^2, i = 0..n-2.

    n = 2, a = 1, b = 100 is the reference's tests/pygradflow/rosenbrock.py:7-46.
    """

    def __init__(self, a, b, lb, ub):
        super().__init__(lb, ub, num_cons=0)
        self.a = np.asarray(a, dtype=np.float64)
        self.b = np.asarray(b, dtype=np.float64)
        assert self.a.shape == (self.num_vars - 1,) and self.b.shape == self.a.shape

    def obj(self, x):
        r = x[1:] - x[:-1] ** 2
        return float(np.sum(self.b * r * r + (self.a - x[:-1]) ** 2))

    def obj_grad(self, x):
        r = x[1:] - x[:-1] ** 2
        g = np.zeros_like(x)
        g[:-1] += -4.0 * self.b * r * x[:-1] - 2.0 * (self.a - x[:-1])
        g[1:] += 2.0 * self.b * r
        return g

    def lag_hess(self, x, y):
        n = x.shape[0]
        r = x[1:] - x[:-1] ** 2
        H = np.zeros((n, n))
        idx = np.arange(n - 1)
        H[idx, idx] += 8.0 * self.b * x[:-1] ** 2 - 4.0 * self.b * r + 2.0
        H[idx + 1, idx + 1] += 2.0 * self.b
        off = -4.0 * self.b * x[:-1]
        H[idx, idx + 1] += off
        H[idx + 1, idx] += off
        return H


class OCP(OracleProblem):
    """Discretised nonlinear optimal-control problem (synthetic family, the reference ships no OCP).

    Variables z = (x_1, u_0, x_2, u_1, ..., x_S, u_{S-1}) stage-interleaved, n = S (nx + nu); x_0 fixed.
    Dynamics c_j = x_{j+1} - x_j - h (A_j x_j + B_j u_j + 0.1 sin(x_j)) = 0, j = 0..S-1  (m = S nx);
    cost sum_j 1/2 (x_{j+1}' Q_j x_{j+1} + u_j' R_j u_j) with diagonal Q, R; |u| <= umax, states free.
    All sums are accumulated sequentially in index order so that the CUDA evaluators reproduce them.
    """

    def __init__(self, A, Bm, Q, R, xinit, umax, h, sparse=False):
        self.sparse = bool(sparse)  # return scipy.sparse J / H like a reference Problem would (problem.py:160-192)
        self.A = np.asarray(A, dtype=np.float64)
        self.Bm = np.asarray(Bm, dtype=np.float64)
        self.Q = np.asarray(Q, dtype=np.float64)
        self.R = np.asarray(R, dtype=np.float64)
        self.xinit = np.asarray(xinit, dtype=np.float64)
        self.h = float(h)
        self.S, self.nx, _ = self.A.shape
        self.nu = self.Bm.shape[2]
        S, nx, nu = self.S, self.nx, self.nu
        lb = np.tile(np.concatenate([np.full(nx, -np.inf), np.full(nu, -float(umax))]), S)
        ub = np.tile(np.concatenate([np.full(nx, np.inf), np.full(nu, float(umax))]), S)
        super().__init__(lb, ub, num_cons=S * nx)

    def _split(self, z):
        Z = np.asarray(z, dtype=np.float64).reshape(self.S, self.nx + self.nu)
        X1, U = Z[:, : self.nx], Z[:, self.nx:]
        Xprev = np.vstack([self.xinit[None, :], X1[:-1]])
        return X1, U, Xprev

    def obj(self, z):
        X1, U, _ = self._split(z)
        return float(0.5 * np.sum(self.Q * X1 * X1) + 0.5 * np.sum(self.R * U * U))

    def obj_grad(self, z):
        X1, U, _ = self._split(z)
        return np.concatenate([self.Q * X1, self.R * U], axis=1).reshape(-1)

    def cons(self, z):
        X1, U, Xprev = self._split(z)
        ax = np.zeros((self.S, self.nx))
        for k in range(self.nx):
            ax = ax + self.A[:, :, k] * Xprev[:, k : k + 1]
        bu = np.zeros((self.S, self.nx))
        for k in range(self.nu):
            bu = bu + self.Bm[:, :, k] * U[:, k : k + 1]
        f = (ax + bu) + 0.1 * np.sin(Xprev)
        return ((X1 - Xprev) - self.h * f).reshape(-1)

    def cons_jac(self, z):
        X1, U, Xprev = self._split(z)
        S, nx, nu = self.S, self.nx, self.nu
        w = nx + nu
        J = np.zeros((S * nx, S * w))
        eye = np.eye(nx)
        for j in range(S):
            rows = slice(j * nx, (j + 1) * nx)
            J[rows, j * w : j * w + nx] = eye
            J[rows, j * w + nx : (j + 1) * w] = -(self.h * self.Bm[j])
            if j >= 1:
                e = self.h * (self.A[j] + eye * (0.1 * np.cos(Xprev[j]))[None, :])
                J[rows, (j - 1) * w : (j - 1) * w + nx] = -(eye + e)
        return scipy.sparse.csr_matrix(J) if self.sparse else J

    def lag_hess(self, z, y):
        X1, U, _ = self._split(z)
        S, nx = self.S, self.nx
        Y = np.asarray(y, dtype=np.float64).reshape(S, nx)
        c1 = 0.1 * self.h
        dx = self.Q.copy()
        dx[:-1] = self.Q[:-1] + (Y[1:] * c1) * np.sin(X1[:-1])
        d = np.concatenate([dx, self.R], axis=1).reshape(-1)
        return scipy.sparse.diags([d], [0], format="csr") if self.sparse else np.diag(d)


class Tame(OracleProblem):
    """f=(x0-x1)^2, c = x0+x1-1 (tests/pygradflow/tame.py:7-36)."""

    def __init__(self):
        super().__init__(np.full(2, -np.inf), np.full(2, np.inf), num_cons=1)

    def obj(self, z):
        return (z[0] - z[1]) ** 2

    def obj_grad(self, z):
        d = z[0] - z[1]
        return np.array([2 * d, -2 * d])

    def cons(self, z):
        return np.array([z[0] + z[1] - 1])

    def cons_jac(self, z):
        return np.array([[1.0, 1.0]])

    def lag_hess(self, z, lag):
        return np.array([[2.0, -2.0], [-2.0, 2.0]])


class HS71(OracleProblem):
    """Hock-Schittkowski 71 with a slack on the product constraint (tests/pygradflow/hs71.py:7-95)."""

    def __init__(self):
        super().__init__(
            np.array([1.0, 1.0, 1.0, 1.0, 0.0]),
            np.array([5.0, 5.0, 5.0, 5.0, np.inf]),
            num_cons=2,
        )

    def obj(self, x):
        return x[0] * x[3] * (x[0] + x[1] + x[2]) + x[2]

    def obj_grad(self, x):
        s = x[0] + x[1] + x[2]
        return np.array([s * x[3] + x[0] * x[3], x[0] * x[3], x[0] * x[3] + 1, s * x[0], 0.0])

    def cons(self, x):
        xx = x[:4]
        return np.array([np.prod(xx) - x[4] - 25.0, np.dot(xx, xx) - 40.0])

    def cons_jac(self, x):
        a, b, c, d = x[:4]
        return np.array(
            [[b * c * d, a * c * d, a * b * d, a * b * c, -1.0], [2 * a, 2 * b, 2 * c, 2 * d, 0.0]]
        )

    def lag_hess(self, x, lag):
        a, b, c, d = x[:4]
        H = np.zeros((5, 5))
        s = 2 * a + b + c
        H[:4, :4] = [[2 * d, d, d, s], [d, 0, 0, a], [d, 0, 0, a], [s, a, a, 0]]
        P = np.array(
            [[0, c * d, b * d, b * c], [c * d, 0, a * d, a * c], [b * d, a * d, 0, a * b], [b * c, a * c, a * b, 0]]
        )
        H[:4, :4] += lag[0] * P + lag[1] * 2.0 * np.eye(4)
        return H


# --------------------------------------------------------------------------
# Iterate (pygradflow/iterate.py:19-208)
# --------------------------------------------------------------------------
def norm_mult(*vs) -> float:
    """pygradflow/util.py:15-24."""
    total = 0.0
    for v in vs:
        total += np.dot(v, v)
    return float(np.sqrt(total))


class Iterate:
    def __init__(self, problem, params, x, y):
        self.problem = problem
        self.params = params
        self.x = np.array(x, dtype=np.float64)
        self.y = np.array(y, dtype=np.float64)
        assert self.x.shape == (problem.num_vars,) and self.y.shape == (problem.num_cons,)
        self._cache = {}

    def _get(self, key, fn):
        if key not in self._cache:
            self._cache[key] = fn()
        return self._cache[key]

    # evaluations: iterate.py:59-76 (cached)
    @property
    def obj(self):
        return self._get("obj", lambda: float(self.problem.obj(self.x)))

    @property
    def obj_grad(self):
        return self._get("g", lambda: np.asarray(self.problem.obj_grad(self.x), dtype=np.float64))

    @property
    def cons(self):
        if self.problem.num_cons == 0:
            return np.zeros(0)
        return self._get("c", lambda: np.asarray(self.problem.cons(self.x), dtype=np.float64))

    @property
    def cons_jac(self):
        if self.problem.num_cons == 0:
            return np.zeros((0, self.problem.num_vars))
        return self._get("J", lambda: self._mat(self.problem.cons_jac(self.x)))

    def _mat(self, a):
        if getattr(self.params, "sparse", False):
            return scipy.sparse.csr_matrix(a)
        return _dense(a)

    def lag_hess(self, y):
        return self._mat(self.problem.lag_hess(self.x, y))

    # iterate.py:91-110
    def aug_lag_deriv_x(self, rho):
        return self.obj_grad + self.cons_jac.T @ (rho * self.cons + self.y)

    def aug_lag_deriv_y(self):
        return self.cons

    def aug_lag_deriv_xy(self):
        return self.cons_jac

    def aug_lag_deriv_xx(self, rho):
        mult = self.y + rho * self.cons
        if rho == 0.0:
            return self.lag_hess(mult)
        J = self.cons_jac
        return self.lag_hess(mult) + rho * (J.T @ J)

    # active_set.py:4-29
    def bound_sets(self):
        def make():
            tol = self.params.active_tol
            lb, ub, x = self.problem.var_lb, self.problem.var_ub, self.x
            lo = np.abs(x - lb) <= tol
            up = np.abs(ub - x) <= tol
            both = lo & up
            return (lo & ~both, up & ~both, both)

        return self._get("sets", make)

    # iterate.py:136-181
    @property
    def bounds_dual(self):
        def make():
            r = -(self.obj_grad + self.cons_jac.T @ self.y)
            at_lower, at_upper, at_both = self.bound_sets()
            d = np.zeros_like(self.x)
            d[at_upper] = np.maximum(r[at_upper], 0.0)
            d[at_lower] = np.minimum(r[at_lower], 0.0)
            d[at_both] = r[at_both]
            return d

        return self._get("d", make)

    @property
    def bound_violation(self):
        lb, ub, x = self.problem.var_lb, self.problem.var_ub, self.x
        if x.size == 0:
            return 0.0
        return max(float(np.max(np.maximum(lb - x, 0.0))), float(np.max(np.maximum(x - ub, 0.0))))

    @property
    def cons_violation(self):
        c = self.cons
        return 0.0 if c.size == 0 else float(np.max(np.abs(c)))

    @property
    def stat_res(self):
        r = self.obj_grad + self.cons_jac.T @ self.y + self.bounds_dual
        return float(np.max(np.abs(r)))

    @property
    def total_res(self):
        return max(self.cons_violation, self.bound_violation, self.stat_res)

    def is_feasible(self, tol):
        return self.cons_violation <= tol and self.bound_violation <= tol

    def locally_infeasible(self, feas_tol, local_infeas_tol):
        """iterate.py:115-134."""
        if self.cons_violation <= feas_tol:
            return False
        r = self.cons_jac.T @ self.cons
        at_lower, at_upper, _ = self.bound_sets()
        r = r.copy()
        r[at_lower] = np.minimum(r[at_lower], 0.0)
        r[at_upper] = np.maximum(r[at_upper], 0.0)
        return bool(np.max(np.abs(r)) <= local_infeas_tol) if r.size else True

    def dist(self, other):
        return norm_mult(self.x - other.x, self.y - other.y)

    def check_eval(self):
        self.obj
        self.obj_grad
        if self.problem.num_cons > 0:
            self.cons
            self.cons_jac


# --------------------------------------------------------------------------
# Residual functions (pygradflow/implicit_func.py)
# --------------------------------------------------------------------------
ACTIVE_SLACK = 1e-8  # implicit_func.py:44


def box_active_set(p, lb, ub):
    """implicit_func.py:21-44."""
    return np.logical_or(p < lb - ACTIVE_SLACK, p > ub + ACTIVE_SLACK)


def box_project(p, lb, ub, active):
    """implicit_func.py:46-60: clip only the entries flagged active."""
    out = np.array(p, copy=True)
    out[active] = np.clip(p[active], lb[active], ub[active])
    return out


class ScaledImplicitFunc:
    """lambda-scaled residual (implicit_func.py:202-294)."""

    def __init__(self, problem, orig_iterate, dt):
        self.problem = problem
        self.orig_iterate = orig_iterate
        self.dt = dt
        self.lamb = 1.0 / dt
        self.lb = self.lamb * problem.var_lb
        self.ub = self.lamb * problem.var_ub
        self.n = problem.num_vars
        self.m = problem.num_cons

    def projection_initial(self, iterate, rho, tau=None):
        x0 = self.orig_iterate.x
        lamb = self.lamb
        if tau is not None:  # :237-244
            lamb = 1.0 / self.dt
            return (
                lamb * (1 - tau * lamb) * iterate.x
                + (tau * lamb * lamb) * x0
                - (tau * lamb) * iterate.aug_lag_deriv_x(rho)
            )
        return lamb * x0 - iterate.aug_lag_deriv_x(rho)  # :246

    def active_set_at_point(self, p):
        return box_active_set(p, self.lb, self.ub)

    def compute_active_set(self, iterate, rho, tau=None):
        return self.active_set_at_point(self.projection_initial(iterate, rho, tau))

    def project(self, p, active):
        return box_project(p, self.lb, self.ub, active)

    def value_at(self, iterate, rho, active_set=None):
        """:219-231 (tau never enters the residual)."""
        lamb = self.lamb
        y0 = self.orig_iterate.y
        p = self.projection_initial(iterate, rho)
        if active_set is None:
            active_set = self.compute_active_set(iterate, rho)
        xval = lamb * iterate.x - self.project(p, active_set)
        yval = -(lamb * iterate.y - (lamb * y0 + iterate.aug_lag_deriv_y()))
        return np.concatenate([xval, yval])

    def deriv_at(self, iterate, rho, active_set=None):
        """:254-294: [[lamb I + P_I H_rho, P_I J'], [-J, lamb I]] as a dense array."""
        if active_set is None:
            active_set = self.compute_active_set(iterate, rho)
        lamb = 1.0 / self.dt
        H = iterate.aug_lag_deriv_xx(rho)
        J = iterate.aug_lag_deriv_xy()
        keep = np.logical_not(active_set).astype(np.float64)[:, None]
        n, m = self.n, self.m
        F = np.zeros((n + m, n + m))
        F[:n, :n] = lamb * np.eye(n) + keep * H
        F[:n, n:] = keep * J.T
        F[n:, :n] = -J
        F[n:, n:] = lamb * np.eye(m)
        return F


class ImplicitFunc:
    """Unscaled residual (implicit_func.py:102-199); used for the controller's ||F||."""

    def __init__(self, problem, orig_iterate, dt):
        self.problem = problem
        self.orig_iterate = orig_iterate
        self.dt = dt
        self.n = problem.num_vars
        self.m = problem.num_cons

    def projection_initial(self, iterate, rho, tau=None):
        x0 = self.orig_iterate.x
        dt = self.dt
        if tau is not None:
            lamb = 1.0 / dt
            return (1.0 - tau * lamb) * iterate.x + (tau * lamb) * x0 - tau * iterate.aug_lag_deriv_x(rho)
        return x0 - dt * iterate.aug_lag_deriv_x(rho)

    def active_set_at_point(self, p):
        return box_active_set(p, self.problem.var_lb, self.problem.var_ub)

    def compute_active_set(self, iterate, rho, tau=None):
        return self.active_set_at_point(self.projection_initial(iterate, rho, tau))

    def project(self, p, active):
        return box_project(p, self.problem.var_lb, self.problem.var_ub, active)

    def value_at(self, iterate, rho, active_set=None):
        """:150-161."""
        y0 = self.orig_iterate.y
        dt = self.dt
        p = self.projection_initial(iterate, rho)
        if active_set is None:
            active_set = self.compute_active_set(iterate, rho)
        xval = iterate.x - self.project(p, active_set)
        yval = iterate.y - (y0 + dt * iterate.aug_lag_deriv_y())
        return np.concatenate([xval, yval])

    def deriv_at(self, iterate, rho, active_set=None):
        """:163-199."""
        if active_set is None:
            active_set = self.compute_active_set(iterate, rho)
        dt = self.dt
        H = iterate.aug_lag_deriv_xx(rho)
        J = iterate.aug_lag_deriv_xy()
        keep = np.logical_not(active_set).astype(np.float64)[:, None]
        n, m = self.n, self.m
        F = np.zeros((n + m, n + m))
        F[:n, :n] = np.eye(n) + keep * (dt * H)
        F[:n, n:] = keep * (dt * J.T)
        F[n:, :n] = -dt * J
        F[n:, n:] = np.eye(m)
        return F


# --------------------------------------------------------------------------
# Linear solver (pygradflow/linear_solver/lu_solver.py:8-21)
# --------------------------------------------------------------------------
class OracleLUSolver:
    def __init__(self, K: np.ndarray, kind: str = "splu", symmetric: bool = False):
        self.K = K
        self.N = K.shape[0]
        self.kind = kind
        self.symmetric = symmetric
        self._neg = None
        if self.N == 0:
            return
        if kind == "splu":
            try:
                self.lu = scipy.sparse.linalg.splu(scipy.sparse.csc_matrix(K))
            except RuntimeError as err:  # exactly singular (lu_solver.py:15-17)
                raise LinearSolverError("LU decomposition failed") from err
        else:
            if not np.all(np.isfinite(K)):
                raise LinearSolverError("non-finite matrix")
            self.lu = scipy.linalg.lu_factor(K, check_finite=False)
            if np.any(np.diag(self.lu[0]) == 0.0):
                raise LinearSolverError("LU decomposition failed")

    def solve(self, rhs, trans=False, initial_sol=None):
        if self.N == 0:
            return np.zeros(0)
        if self.kind == "splu":
            return self.lu.solve(rhs, trans="T" if trans else "N")
        return scipy.linalg.lu_solve(self.lu, rhs, trans=1 if trans else 0, check_finite=False)

    def num_neg_eigvals(self):
        """The reference LUSolver returns None; the oracle offers the exact count for the
        inertia contract of the symmetric solvers (ma57_solver.py:76-79 et al.)."""
        if not self.symmetric:
            return None
        if self._neg is None:
            self._neg = 0 if self.N == 0 else int((np.linalg.eigvalsh(0.5 * (self.K + self.K.T)) < 0).sum())
        return self._neg

    def rcond(self):
        return None


class OracleIterativeSolver:
    """linear_solver/gmres_solver.py:7-35 and linear_solver/minres_solver.py:6-24: the wrappers restated line by line
    around the SAME third-party routines (scipy.sparse.linalg.gmres / minres, SciPy 1.18.1 here; the reference leaves
    scipy unpinned).  ``matvecs`` counts the products of the last solve (the reference does not expose it; the CUDA
    kernels report the same count)."""

    def __init__(self, K, kind: str, symmetric: bool = False):
        assert kind in ("gmres", "minres")
        if kind == "minres":
            assert symmetric, "MINRES requires a symmetric matrix"     # minres_solver.py:9
        self.kind = kind
        self.mat = scipy.sparse.csc_matrix(K)
        self.N = self.mat.shape[0]
        self.symmetric = symmetric
        self.matvecs = 0

    def _counted(self, mat):
        def mv(v):
            self.matvecs += 1
            return mat @ v

        return scipy.sparse.linalg.LinearOperator(mat.shape, matvec=mv, dtype=np.float64)

    def solve(self, rhs, trans=False, initial_sol=None):
        self.matvecs = 0
        if self.N == 0:
            return np.zeros(0)
        if self.kind == "minres":
            if initial_sol is not None:
                initial_sol = initial_sol()
            sol, info = scipy.sparse.linalg.minres(self._counted(self.mat), rhs, x0=initial_sol)
            if info != 0:
                raise LinearSolverError("MINRES failed with error code {}".format(info))
            return sol
        mat = self.mat.T if trans else self.mat
        if initial_sol is not None:
            initial_sol = initial_sol()
        n = mat.shape[0]
        atol = 1e-8
        if initial_sol is not None:                                      # gmres_solver.py:22-25
            res = rhs - mat @ initial_sol
            if np.linalg.norm(res, ord=np.inf) < atol:
                return initial_sol
        sol, info = scipy.sparse.linalg.gmres(self._counted(mat), rhs, maxiter=n, x0=initial_sol, atol=atol)
        if info != 0:
            raise LinearSolverError("GMRES failed with error code {}".format(info))
        return sol

    def num_neg_eigvals(self):
        return None

    def rcond(self):
        return None


def make_linear_solver(K, kind: str, symmetric: bool):
    """linear_solver/__init__.py:8-39."""
    if kind in ("gmres", "minres"):
        return OracleIterativeSolver(K, kind, symmetric=symmetric)
    return OracleLUSolver(K, kind, symmetric=symmetric)


class ConditionEstimator:
    """step/cond_estimate.py:13-114 (Dixon): 1/cond_2 from power iterations on A'A and (A'A)^-1, the latter through
    solve(trans=True) / solve of the factorised matrix; random start vectors from default_rng(42)."""

    def __init__(self, mat, linear_solver, min_prob=0.99, factor=10.0):
        self.mat = np.asarray(mat, dtype=np.float64)
        self.size = self.mat.shape[0]
        self.linear_solver = linear_solver
        self.min_prob, self.factor = min_prob, factor
        self.rng = np.random.default_rng(seed=42)

    def required_its(self):                                            # :41-43
        f = (1.0 - self.min_prob) / 1.6 * math.pow(self.size, -0.5)
        return -2 * math.ceil(math.log(f, self.factor))

    def _random_vec(self):                                             # :45-57
        vec = self.rng.normal(size=self.size)
        while not (vec != 0.0).any():
            vec = self.rng.normal(size=self.size)
        return vec / np.linalg.norm(vec)

    def estimate_rcond(self):                                          # :59-114
        mat, ls = self.mat, self.linear_solver
        num_its = self.required_its()
        x, y = self._random_vec(), self._random_vec()
        xprod, yprod = np.copy(x), np.copy(y)
        xfac = yfac = 1.0
        for _ in range(num_its):
            xprod = mat.T @ (mat @ xprod)
            yprod = ls.solve(ls.solve(yprod, trans=True))
            xnorm, ynorm = float(np.linalg.norm(xprod)), float(np.linalg.norm(yprod))
            xfac *= xnorm
            xprod /= xnorm
            yfac *= ynorm
            yprod /= ynorm
        pow_fac = 1.0 / (2.0 * num_its)
        xdot = math.pow(x.dot(xprod) * xfac, pow_fac)
        ydot = math.pow(y.dot(yprod) * yfac, pow_fac)
        if np.isinf(xdot) or np.isinf(ydot) or np.isinf(xdot * ydot):
            return 0.0
        return 1.0 / (xdot * ydot)


# --------------------------------------------------------------------------
# Step result / step solver (pygradflow/step/solver/*.py)
# --------------------------------------------------------------------------
class StepResult:
    """step_solver.py:16-63."""

    def __init__(self, orig_iterate, dx, dy, active_set, rcond=None):
        self.orig_iterate = orig_iterate
        self.dy = dy
        self.active_set = active_set
        self.rcond = rcond
        lb, ub = orig_iterate.problem.var_lb, orig_iterate.problem.var_ub
        x = orig_iterate.x
        xn = x - dx
        dx = np.array(dx, copy=True)
        low = xn < lb
        xn[low] = lb[low]
        dx[low] = x[low] - lb[low]
        high = xn > ub
        xn[high] = ub[high]
        dx[high] = x[high] - ub[high]
        self.dx = dx
        self.xn = xn
        self._iterate = None

    @property
    def iterate(self):
        if self._iterate is None:
            o = self.orig_iterate
            self._iterate = Iterate(o.problem, o.params, self.xn, o.y - self.dy)
        return self._iterate

    @property
    def diff(self):
        return norm_mult(self.dx, self.dy)


def kkt_system(H0, J, active, lamb, rho, b0, b1, b2t):
    """Dense reduced symmetric KKT matrix and right-hand side.

    symmetric_step_solver.py:27-39 (H0 + lamb I, inactive rows), :49-77 (bmat), :79-94 (rhs).
    Returns (K, rhs) with K of order |I| + m.
    """
    inactive = np.logical_not(active)
    n = H0.shape[0]
    m = J.shape[0]
    Hl = H0 + lamb * np.eye(n)
    HI = Hl[inactive, :]
    JI = J[:, inactive]
    nI = int(inactive.sum())
    K = np.zeros((nI + m, nI + m))
    K[:nI, :nI] = HI[:, inactive]
    K[:nI, nI:] = JI.T
    K[nI:, :nI] = JI
    K[nI:, nI:] = (-lamb / (1.0 + lamb * rho)) * np.eye(m)
    rhs = np.concatenate([b1 - HI[:, active] @ b0, b2t - J[:, active] @ b0])
    return K, rhs


def kkt_system_sparse(H0, J, active, lamb, rho, b0, b1, b2t):
    """kkt_system with scipy.sparse operands, statement by statement as the reference builds it
    (symmetric_step_solver.py:27-39 compute_hess_jac, :49-77 _compute_deriv, :79-94 compute_rhs)."""
    inactive_indices = np.where(np.logical_not(active))[0]
    active_indices = np.where(active)[0]
    n = H0.shape[0]
    m = J.shape[0]
    hess = H0 + scipy.sparse.diags([lamb], shape=(n, n), dtype=np.float64)
    hess_rows = hess.tocsr()[inactive_indices, :].tocsc()
    jac = J.tocsc()
    inactive_jac = jac[:, inactive_indices]
    inactive_hess = hess_rows[:, inactive_indices]
    lower = scipy.sparse.diags([-lamb / (1.0 + lamb * rho)], shape=(m, m), dtype=np.float64)
    K = scipy.sparse.bmat([[inactive_hess, inactive_jac.T], [inactive_jac, lower]], format="csc")
    rhs = np.concatenate((b1 - (hess_rows[:, active_indices] @ b0), b2t - (jac[:, active_indices] @ b0)))
    return K, rhs


class SymmetricStepSolver:
    """scaled_step_solver.py:15-107 + symmetric_step_solver.py:13-164 on dense arrays."""

    def __init__(self, problem, params, orig_iterate, dt, rho):
        assert dt > 0.0 and rho > 0.0
        self.problem = problem
        self.params = params
        self.orig_iterate = orig_iterate
        self.dt = dt
        self.rho = rho
        self.n = problem.num_vars
        self.m = problem.num_cons
        self._func = ScaledImplicitFunc(problem, orig_iterate, dt)
        self.active_set = None
        self.jac = None
        self.hess = None
        self.solver = None
        self.K = None
        self.num_factorizations = 0

    @property
    def func(self):
        return self._func

    def update_derivs(self, iterate):
        J, H = iterate.aug_lag_deriv_xy(), iterate.aug_lag_deriv_xx(0.0)  # multiplier y only
        self.jac = J.copy() if scipy.sparse.issparse(J) else np.array(J, copy=True)
        self.hess = H.copy() if scipy.sparse.issparse(H) else np.array(H, copy=True)
        self.solver = None
        self.K = None

    def update_active_set(self, active_set):
        self.active_set = np.array(active_set, copy=True)
        self.solver = None
        self.K = None

    def initial_rhs(self, iterate):
        """scaled_step_solver.py:38-60."""
        F = self._func.value_at(iterate, self.rho, self.active_set)
        rx, ry = F[: self.n], F[self.n :]
        A = self.active_set
        return self.dt * rx[A], rx[~A], ry

    def linear_solver(self, K):
        return make_linear_solver(K, self.params.linear_solver, symmetric=True)

    def solve(self, iterate):
        """scaled_step_solver.py:85-107 + symmetric_step_solver.py:96-164."""
        b0, b1, b2 = self.initial_rhs(iterate)
        rho = self.rho
        lamb = 1.0 / self.dt
        fact = 1.0 / (1.0 + lamb * rho)
        b2t = fact * b2
        A = self.active_set
        build = kkt_system_sparse if scipy.sparse.issparse(self.hess) else kkt_system
        K, rhs = build(self.hess, self.jac, A, lamb, rho, b0, b1, b2t)
        if self.K is None:
            self.K = K
        try:
            if self.solver is None:
                self.solver = self.linear_solver(self.K)
                self.num_factorizations += 1
            s = self.solver.solve(rhs)
            if self.params.inertia_correction:
                neg = self.solver.num_neg_eigvals()
                if neg is None:
                    raise Exception("Inertia correction requested but not available")
                if neg != self.m:
                    raise LinearSolverError("Invalid matrix inertia")
        except LinearSolverError as err:
            raise StepSolverError() from err
        self.last_rhs = rhs
        self.last_sol = s
        nI = int((~A).sum())
        dx = np.zeros(self.n)
        dx[~A] = s[:nI]
        dx[A] = b0
        dy = fact * (s[nI:] - rho * b2)
        return StepResult(iterate, dx, dy, self.active_set, None)


class AsymmetricStepSolver(SymmetricStepSolver):
    """asymmetric_step_solver.py:15-173: the full (n+m) system [[H + lamb I, J'], [J, -lamb fact I]] in the natural
    order, the row of every active variable overwritten by the unit row (:37-75), rhs = (b0 | b1 by position, b2t)
    (:106-123).  Same ScaledStepSolver shell as the symmetric solver (scaled_step_solver.py:85-107)."""

    def full_system(self, b0, b1, b2t):
        n, m, A = self.n, self.m, self.active_set
        lamb = 1.0 / self.dt
        K = np.zeros((n + m, n + m))
        K[:n, :n] = self.hess + lamb * np.eye(n)
        K[:n, n:] = self.jac.T
        K[n:, :n] = self.jac
        K[n:, n:] = (-lamb / (1.0 + lamb * self.rho)) * np.eye(m)
        K[:n][A] = 0.0
        K[np.where(A)[0], np.where(A)[0]] = 1.0
        rhs = np.empty(n + m)
        rhs[n:] = b2t
        rhs[:n][A] = b0
        rhs[:n][~A] = b1
        return K, rhs

    def linear_solver(self, K):
        return make_linear_solver(K, self.params.linear_solver, symmetric=False)

    def initial_sol(self, b0):
        """asymmetric_step_solver.py:125-138 (None for the Extended solver, which passes none)."""
        n, m, A = self.n, self.m, self.active_set

        def initial_sol():
            sol = np.zeros(n + m)
            sol[:n][A] = b0
            return sol

        return initial_sol

    def solve(self, iterate):
        b0, b1, b2 = self.initial_rhs(iterate)
        rho = self.rho
        lamb = 1.0 / self.dt
        fact = 1.0 / (1.0 + lamb * rho)
        b2t = fact * b2
        K, rhs = self.full_system(b0, b1, b2t)
        if self.K is None:
            self.K = K
        try:
            if self.solver is None:
                self.solver = self.linear_solver(self.K)
                self.num_factorizations += 1
            s = self.solver.solve(rhs, initial_sol=self.initial_sol(b0))
        except LinearSolverError as err:
            raise StepSolverError() from err
        self.last_rhs = rhs
        self.last_sol = s
        dx = s[: self.n]
        dy = fact * (s[self.n :] - rho * b2)
        return StepResult(iterate, dx, dy, self.active_set, None)


class ExtendedStepSolver(AsymmetricStepSolver):
    """extended_step_solver.py:12-112: rows = (selector rows of the active variables; inactive rows of
    [H + lamb I, J']; [J, -lamb fact I]), columns in the natural order, rhs = (b0, b1, b2t) (:85-95)."""

    def full_system(self, b0, b1, b2t):
        n, m, A = self.n, self.m, self.active_set
        lamb = 1.0 / self.dt
        act, ina = np.where(A)[0], np.where(~A)[0]
        nA = act.size
        K = np.zeros((n + m, n + m))
        K[np.arange(nA), act] = 1.0
        K[nA:n, :n] = (self.hess + lamb * np.eye(n))[ina, :]
        K[nA:n, n:] = self.jac.T[ina, :]
        K[n:, :n] = self.jac
        K[n:, n:] = (-lamb / (1.0 + lamb * self.rho)) * np.eye(m)
        return K, np.concatenate([b0, b1, b2t])

    def initial_sol(self, b0):
        return None                                                      # extended_step_solver.py:98: solve(rhs) only


class StandardStepSolver:
    """standard_step_solver.py:15-92: Newton step on the UNSCALED implicit function (implicit_func.py:100-199);
    matrix F'(x_hat) with H_rho = H + rho J'J (:50-53), rhs = F(iterate) (:63), dx = sol[:n], dy = sol[n:]."""

    def __init__(self, problem, params, orig_iterate, dt, rho):
        self.problem = problem
        self.params = params
        self.orig_iterate = orig_iterate
        self.dt = dt
        self.rho = rho
        self.n = problem.num_vars
        self.m = problem.num_cons
        self._func = ImplicitFunc(problem, orig_iterate, dt)
        self.active_set = None
        self.jac = None
        self.hess = None
        self.solver = None
        self.K = None
        self.num_factorizations = 0

    @property
    def func(self):
        return self._func

    def update_derivs(self, iterate):
        self.jac = np.array(iterate.aug_lag_deriv_xy(), copy=True)
        self.hess = np.array(iterate.aug_lag_deriv_xx(self.rho), copy=True)
        self.solver = None
        self.K = None

    def update_active_set(self, active_set):
        self.active_set = np.array(active_set, copy=True)
        self.solver = None
        self.K = None

    def matrix(self):
        """implicit_func.py:163-187 with the frozen derivatives."""
        n, m, dt = self.n, self.m, self.dt
        keep = np.logical_not(self.active_set).astype(np.float64)[:, None]
        K = np.zeros((n + m, n + m))
        K[:n, :n] = np.eye(n) + keep * (dt * self.hess)
        K[:n, n:] = keep * (dt * self.jac.T)
        K[n:, :n] = -dt * self.jac
        K[n:, n:] = np.eye(m)
        return K

    def solve(self, iterate):
        n, m = self.n, self.m
        if self.K is None:
            self.K = self.matrix()
        rhs = self._func.value_at(iterate, self.rho, self.active_set)
        try:
            if self.solver is None:
                self.solver = make_linear_solver(self.K, self.params.linear_solver, symmetric=False)
                self.num_factorizations += 1
            s = self.solver.solve(rhs)
        except LinearSolverError as err:
            raise StepSolverError() from err
        self.last_rhs = rhs
        self.last_sol = s
        return StepResult(iterate, s[:n], s[n:], self.active_set, None)


def make_step_solver(problem, params, iterate, dt, rho):
    """step/solver/__init__.py:12-31."""
    assert dt > 0.0 and rho > 0.0
    if params.step_solver is not None:
        return params.step_solver(problem, params, iterate, dt, rho)
    cls = {
        "symmetric": SymmetricStepSolver,
        "asymmetric": AsymmetricStepSolver,
        "extended": ExtendedStepSolver,
        "standard": StandardStepSolver,
    }[params.step_solver_type]
    return cls(problem, params, iterate, dt, rho)


# --------------------------------------------------------------------------
# Newton methods (pygradflow/newton.py)
# --------------------------------------------------------------------------
class NewtonMethod:
    def __init__(self, problem, orig_iterate, dt, rho, solver, tau=None):
        self.problem = problem
        self.orig_iterate = orig_iterate
        self.dt = dt
        self.rho = rho
        self.tau = tau
        self.step_solver = solver
        self.func = solver.func


class SimplifiedNewton(NewtonMethod):
    """newton.py:35-60: active set and derivatives frozen at the initial iterate."""

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        A = self.func.compute_active_set(self.orig_iterate, self.rho, self.tau)
        self.step_solver.update_active_set(A)
        self.step_solver.update_derivs(self.orig_iterate)

    def step(self, iterate):
        return self.step_solver.solve(iterate)


class FullNewton(NewtonMethod):
    """newton.py:63-89."""

    def step(self, iterate):
        A = self.func.compute_active_set(iterate, self.rho, self.tau)
        self.step_solver.update_active_set(A)
        self.step_solver.update_derivs(iterate)
        return self.step_solver.solve(iterate)


class ActiveSetNewton(NewtonMethod):
    """newton.py:181-215: derivatives frozen, refactor only when the active set changes."""

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self.step_solver.update_derivs(self.orig_iterate)
        self._current = None

    def step(self, iterate):
        A = self.func.compute_active_set(iterate, self.rho, self.tau)
        if self._current is None or (self._current != A).any():
            self.step_solver.update_active_set(A)
        self._current = A
        return self.step_solver.solve(iterate)


class GlobalizedNewton(NewtonMethod):
    """newton.py:218-304 including its quirks (rhs at orig_iterate, '+' Armijo sign)."""

    def step(self, iterate):
        params = iterate.params
        self.step_solver.update_derivs(iterate)
        A = self.func.compute_active_set(iterate, self.rho, self.tau)
        self.step_solver.update_active_set(A)
        result = self.step_solver.solve(self.orig_iterate)  # :248
        F = self.func.value_at(iterate, self.rho)
        res = 0.5 * np.dot(F, F)
        if res <= params.newton_tol:
            return result
        grad = self.func.deriv_at(iterate, self.rho).T @ F
        n = self.problem.num_vars
        dx, dy = result.dx, result.dy
        inner = np.dot(grad[:n], dx) + np.dot(grad[n:], dy)
        alpha = 1.0
        self.trials = 0
        for _ in range(30):
            self.trials += 1
            trial = Iterate(self.problem, params, iterate.x - dx, iterate.y - dy)
            Fn = self.func.value_at(trial, self.rho)
            nres = 0.5 * np.dot(Fn, Fn)
            if nres <= params.newton_tol:
                break
            if nres <= res + (1e-4 * alpha * inner):
                break
            alpha *= 0.5
            dx = alpha * result.dx
            dy = alpha * result.dy
        else:
            raise LineSearchError("Line search failed to converge")
        out = StepResult(self.orig_iterate, dx, dy, None, None)
        out.active_set = self.func.compute_active_set(out.iterate, self.rho, self.tau)
        return out


def newton_method(problem, params, iterate, dt, rho, tau=None):
    """newton.py:307-323."""
    assert dt > 0.0 and rho > 0.0
    solver = make_step_solver(problem, params, iterate, dt, rho)
    cls = {
        "simplified": SimplifiedNewton,
        "full": FullNewton,
        "active_set": ActiveSetNewton,
        "globalized": GlobalizedNewton,
    }[params.newton_type]
    return cls(problem, iterate, dt, rho, solver, tau)


# --------------------------------------------------------------------------
# Step-size control (distance_ratio_control.py, controller.py, step_control.py)
# --------------------------------------------------------------------------
class LogPIController:
    """controller.py:29-77.  The integral term is never reset."""

    def __init__(self, params):
        self.K_P = params.K_P
        self.K_I = params.K_I
        self.log_ref = math.log(params.theta_ref)
        self.error_sum = 0.0

    def update(self, theta):
        assert theta > 0.0
        err = self.log_ref - math.log(theta)
        self.error_sum += err
        return math.exp(self.K_P * err + self.K_I * self.error_sum)


@dataclass
class ControlResult:
    iterate: Iterate
    lamb: float
    active_set: Optional[np.ndarray]
    accepted: bool
    newton_steps: int = 0
    theta: float = float("nan")


class DistanceRatioController:
    """distance_ratio_control.py:12-78; one instance lives for the whole solve."""

    def __init__(self, problem, params):
        self.problem = problem
        self.params = params
        self.pi = LogPIController(params)
        self.total_newton_steps = 0
        self.total_factorizations = 0
        self.trace_hook = None

    def _newton(self, iterate, rho, dt):
        # newton_control.py:22-38
        return newton_method(self.problem, self.params, iterate, dt, rho, self.compute_tau(iterate, rho))

    def tau_vals(self, iterate, rho):
        """newton_control.py:40-58."""
        x, g = iterate.x, iterate.aug_lag_deriv_x(rho)
        xl, xu = self.problem.var_lb, self.problem.var_ub
        nonzero = np.logical_not(np.isclose(g, 0.0))
        pos, neg = (g > 0.0) & nonzero, (g < 0.0) & nonzero
        tv = np.full_like(x, fill_value=-1)
        tv[pos] = (x[pos] - xl[pos]) / g[pos]
        tv[neg] = (xu[neg] - x[neg]) / -g[neg]
        return tv

    def compute_tau(self, iterate, rho):
        """newton_control.py:60-88 (no user active_set_method)."""
        t = self.params.active_set_type
        if t == "explicit":
            assert self.params.active_set_tau is not None
            return self.params.active_set_tau
        if t == "standard":
            return None
        tv = self.tau_vals(iterate, rho)
        if t == "smallest":
            if (tv <= 0).all():
                return 1.0
            return 0.5 * np.min(tv[tv > 0])
        return max(np.max(tv), 1.0)

    def step(self, iterate, rho, dt):
        assert dt > 0.0
        p = self.params
        lamb = 1.0 / dt
        method = self._newton(iterate, rho, dt)
        func = ImplicitFunc(self.problem, iterate, dt)

        mid = method.step(iterate)
        self.total_newton_steps += 1
        if self.trace_hook:
            self.trace_hook(method, 0, mid)
        mid_norm = float(np.linalg.norm(func.value_at(mid.iterate, rho)))
        if mid_norm <= p.newton_tol:
            return ControlResult(mid.iterate, max(lamb * p.lamb_red, p.lamb_min), mid.active_set, True, 1)
        d1 = mid.diff
        if d1 == 0.0:
            return ControlResult(mid.iterate, lamb, mid.active_set, True, 1)

        fin = method.step(mid.iterate)
        self.total_newton_steps += 1
        if self.trace_hook:
            self.trace_hook(method, 1, fin)
        d2 = fin.diff
        if d2 == 0.0:
            return ControlResult(fin.iterate, lamb, fin.active_set, True, 2)
        theta = d2 / d1
        accepted = theta <= p.theta_max
        if accepted:
            lamb_n = max(p.lamb_min, lamb / self.pi.update(theta))
        else:
            lamb_n = lamb * p.lamb_inc
        return ControlResult(fin.iterate, lamb_n, fin.active_set, accepted, 2, theta)

    def compute_step(self, iterate, rho, dt):
        """step_control.py:67-107: solver failure => reject and double lambda."""
        try:
            res = self.step(iterate, rho, dt)
            if res.accepted:
                res.iterate.check_eval()
            return res
        except StepSolverError:
            return ControlResult(iterate, 2.0 * (1.0 / dt), None, False, 0)


class ResiduumRatioController(DistanceRatioController):
    """residuum_ratio_control.py:12-63: one Newton step, theta = |F(mid)| / |F(orig)| (unscaled residual)."""

    def step(self, iterate, rho, dt):
        p = self.params
        lamb = 1.0 / dt
        method = self._newton(iterate, rho, dt)
        func = ImplicitFunc(self.problem, iterate, dt)
        mid = method.step(iterate)
        self.total_newton_steps += 1
        mid_norm = float(np.linalg.norm(func.value_at(mid.iterate, rho)))
        if mid_norm <= p.newton_tol:
            return ControlResult(mid.iterate, max(lamb * p.lamb_red, p.lamb_min), mid.active_set, True, 1)
        orig_norm = float(np.linalg.norm(func.value_at(iterate, rho)))
        theta = mid_norm / orig_norm
        accepted = theta <= p.theta_max
        if accepted:
            lamb_n = max(p.lamb_min, lamb / self.pi.update(theta))
        else:
            lamb_n = lamb * p.lamb_inc  # (the reference's reset branch is dead: LogController.error_sum stays 0.0)
        return ControlResult(mid.iterate, lamb_n, mid.active_set, accepted, 1, theta)


class ExactController(DistanceRatioController):
    """exact_control.py:10-66: Newton steps until |F| <= newton_tol (accept, lambda / 2) or the contraction rate
    exceeds 1/2 / ten steps are used up (reject, 2 lambda)."""

    max_num_it = 10
    rate_bound = 0.5

    def step(self, iterate, rho, dt):
        p = self.params
        lamb = 1.0 / dt
        func = ImplicitFunc(self.problem, iterate, dt)
        curr = float(np.linalg.norm(func.value_at(iterate, rho)))
        method = self._newton(iterate, rho, dt)
        cur_it = iterate
        for i in range(self.max_num_it):
            st = method.step(cur_it)
            self.total_newton_steps += 1
            nxt = st.iterate
            val = float(np.linalg.norm(func.value_at(nxt, rho)))
            if val <= p.newton_tol:
                return ControlResult(nxt, 0.5 * lamb, st.active_set, True, i + 1)
            if val / curr > self.rate_bound:
                break
            curr = val
            cur_it = nxt
        return ControlResult(nxt, 2.0 * lamb, st.active_set, False, i + 1)


class FixedStepSizeController(DistanceRatioController):
    """fixed_control.py:7-19: one Newton step, always accepted, lambda stays lamb_init."""

    def step(self, iterate, rho, dt):
        st = self._newton(iterate, rho, dt).step(iterate)
        self.total_newton_steps += 1
        return ControlResult(st.iterate, self.params.lamb_init, st.active_set, True, 1)


def step_controller(problem, params):
    """step_control.py:123-150 (the Newton-based controllers)."""
    return {"distance_ratio": DistanceRatioController, "residuum_ratio": ResiduumRatioController,
            "exact": ExactController, "fixed": FixedStepSizeController}[params.step_control_type](problem, params)


# --------------------------------------------------------------------------
# Penalty (pygradflow/penalty.py:36-74)
# --------------------------------------------------------------------------
class Penalty:
    """penalty.py:12-275: ``update(next_iterate)`` returns (next_rho, accept).  ``self.rho`` is the strategy's own
    state; the solver only takes it over when the step is accepted (solver.py:357-369), so after a filter rejection
    (penalty.py:209-210: rho *= 10, reject) the two differ until the next accepted step."""

    def __init__(self, problem, params):
        self.problem = problem
        self.params = params
        self.rho = params.rho
        self.entries = []                                               # PenaltyFilter.entries (penalty.py:176)

    def filter_insert(self, first, second):                            # penalty.py:179-192
        entry = (first, second)

        def dominates(a, b):
            return a[0] <= b[0] and a[1] <= b[1]

        if any(dominates(e, entry) for e in self.entries):
            return False
        self.entries = [e for e in self.entries if not dominates(entry, e)]
        self.entries.append(entry)
        return True

    def update(self, next_iterate):
        kind = self.params.penalty_update
        if kind == "constant":
            return self.params.rho, True
        if kind == "dual_equilibration":                                # penalty.py:77-113
            cons = next_iterate.cons
            yprod = abs(np.dot(next_iterate.y, cons))
            viol = 1.0 / 2.0 * np.dot(cons, cons)
            if viol == 0.0:
                return self.rho, True
            target_rho = 0.01 * yprod / viol
            if self.rho < target_rho:
                self.rho = max(self.rho * 10.0, target_rho)
            return self.rho, True
        if kind == "pareto_decrease":                                   # penalty.py:115-168
            it, prm = next_iterate, self.params
            cons = it.cons
            viol = 1.0 / 2.0 * np.dot(cons, cons)
            if viol <= prm.opt_tol:
                return self.rho, True
            J = it.cons_jac
            infeas_opt_res = J.T.dot(cons)
            if np.linalg.norm(infeas_opt_res, ord=np.inf) <= prm.local_infeas_tol:
                return self.rho, True
            obj_bound = np.inf
            g = it.obj_grad
            obj_prod = np.dot(g, infeas_opt_res)
            cons_dual_prod = J.T.dot(it.y)
            if abs(obj_prod) > 1e-10:
                lhs = -(np.linalg.norm(g) + cons_dual_prod.dot(g))
                obj_bound = lhs / obj_prod
            lhs = -np.dot(infeas_opt_res, g + cons_dual_prod)
            cons_bound = lhs / np.linalg.norm(infeas_opt_res)
            bound = min(obj_bound, cons_bound)
            assert np.isfinite(bound)
            next_rho = max(min(self.rho * 10.0, bound), self.rho)
            self.rho = next_rho
            return self.rho, True
        if kind in ("objective_filter", "lagrangian_filter"):           # penalty.py:170-255
            it = next_iterate
            if kind == "objective_filter":
                entry = (it.obj, it.cons_violation)
            else:
                lag_x = it.aug_lag_deriv_x(self.rho)
                lag_y = it.aug_lag_deriv_y()
                entry = (np.dot(lag_x, lag_x) + np.dot(lag_y, lag_y), float(np.linalg.norm(it.cons)))
            if self.filter_insert(*entry):
                return self.rho, True
            self.rho *= 10.0
            return self.rho, False
        assert kind == "dual_norm", kind
        if self.problem.num_cons == 0:                                  # penalty.py:46-74
            return self.rho, True
        ynorm = float(np.max(np.abs(next_iterate.y)))
        if ynorm >= 10.0 * self.rho:
            self.rho = min(ynorm, 10.0 * self.rho)
        return self.rho, True


# --------------------------------------------------------------------------
# Outer loop (pygradflow/solver.py)
# --------------------------------------------------------------------------
@dataclass
class SolveResult:
    x: np.ndarray
    y: np.ndarray
    d: np.ndarray
    status: int
    iterations: int
    accepted_steps: int
    newton_steps: int
    lamb: float
    rho: float
    trace: List[dict] = field(default_factory=list)

    @property
    def success(self):
        return self.status == STATUS_OPTIMAL


def check_terminate(iterate, iteration, params):
    """solver.py:180-205 (time limit omitted: never set on this path)."""
    if params.iteration_limit is not None and iteration >= params.iteration_limit:
        return STATUS_ITERATION_LIMIT
    if iterate.total_res <= params.opt_tol:
        return STATUS_OPTIMAL
    if iterate.locally_infeasible(params.opt_tol, params.local_infeas_tol):
        return STATUS_LOCALLY_INFEASIBLE
    if iterate.obj <= params.obj_lower_limit and iterate.is_feasible(params.opt_tol):
        return STATUS_UNBOUNDED
    return None


def initial_iterate(problem, params, x0=None, y0=None):
    """transform.py:29-54 without scaling."""
    if x0 is None:
        x = np.clip(np.zeros(problem.num_vars), problem.var_lb, problem.var_ub)
    else:
        x = np.broadcast_to(x0, (problem.num_vars,))
    y = np.zeros(problem.num_cons) if y0 is None else np.broadcast_to(y0, (problem.num_cons,))
    return Iterate(problem, params, x.astype(np.float64), y.astype(np.float64))


class Solver:
    """Scalar restatement of solver.py:26-431 (display, timers, scaling, slack transform omitted)."""

    def __init__(self, problem, params=None):
        self.problem = problem
        self.params = params if params is not None else OracleParams()

    def perform_iteration(self, x0=None, y0=None):
        """solver.py:207-231."""
        p = self.params
        it = initial_iterate(self.problem, p, x0, y0)
        ctrl = step_controller(self.problem, p)
        res = ctrl.compute_step(it, p.rho, 1.0 / p.lamb_init)
        nxt = res.iterate
        return nxt.x, nxt.y, nxt.bounds_dual

    def solve(self, x0=None, y0=None, record=False, step_hook=None) -> SolveResult:
        """solver.py:233-431."""
        p = self.params
        problem = self.problem
        iterate = initial_iterate(problem, p, x0, y0)
        iterate.check_eval()
        lamb = p.lamb_init
        ctrl = step_controller(problem, p)
        ctrl.trace_hook = step_hook
        penalty = Penalty(problem, p)
        rho = penalty.rho
        iteration = 0
        accepted_steps = 0
        trace = []
        while True:
            status = check_terminate(iterate, iteration, p)
            if status is not None:
                break
            res = ctrl.compute_step(iterate, rho, 1.0 / lamb)
            accept = res.accepted
            lamb_used = lamb
            lamb = res.lamb
            if lamb >= p.lamb_max:
                raise LambMaxError(f"Inverse step size {lamb} exceeded maximum {p.lamb_max}")
            if record:
                trace.append(
                    dict(
                        x=res.iterate.x.copy(),
                        y=res.iterate.y.copy(),
                        accept=bool(accept),
                        lamb_used=lamb_used,
                        lamb_next=lamb,
                        rho=rho,
                        active=None if res.active_set is None else res.active_set.copy(),
                        newton_steps=res.newton_steps,
                        theta=res.theta,
                    )
                )
            if accept:                                                  # solver.py:357-378
                next_rho, accept = penalty.update(res.iterate)
                if record:
                    trace[-1]["penalty_accept"] = bool(accept)
            if accept:
                rho = next_rho
                iterate = res.iterate
                accepted_steps += 1
            iteration += 1
        return SolveResult(
            x=iterate.x,
            y=iterate.y,
            d=iterate.bounds_dual,
            status=status,
            iterations=iteration,
            accepted_steps=accepted_steps,
            newton_steps=ctrl.total_newton_steps,
            lamb=lamb,
            rho=rho,
            trace=trace,
        )
