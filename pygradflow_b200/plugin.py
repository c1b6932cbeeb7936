"""Drop-in step solver / linear solver for the reference's plug-in points (batch = 1 views of the
batched CUDA core).

    Solver(problem, Params(step_solver=B200StepSolver)).solve(x0, y0)

``Params.step_solver`` is invoked as ``step_solver(problem, params, iterate, dt, rho)``
(pygradflow/step/solver/__init__.py:18-19, pygradflow/params.py:234) and must return an object with the
``StepSolver`` interface (pygradflow/step/solver/step_solver.py:66-130): property ``func`` and methods
``update_active_set``, ``update_derivs``, ``solve`` -> ``StepResult``.  ``B200LinearSolver`` has the
``LinearSolver`` interface (pygradflow/linear_solver/linear_solver.py:18-31).  Host arrays go in and out;
every floating-point operation of the path runs in the CUDA kernels of libgradflow_b200.so -- problem
callbacks (``Problem.obj_grad`` etc.) stay the user's Python code, exactly as in the reference.
"""

from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from . import kernels as K
from .engine import KKTEngine
from .kernels import WorkList
from .params import LinearSolverType


class LinearSolverError(Exception):
    """Same role as pygradflow/linear_solver/linear_solver.py:8-15."""


class StepSolverError(Exception):
    """Same role as pygradflow/step/step_solver_error.py:1-7."""


_ERROR_TYPES = None


def set_error_types(linear_solver_error, step_solver_error) -> None:
    """Exception classes the plug-in raises.  Under the reference's Solver they are found automatically
    (``_reference_errors``); a host loop with its own classes (the test oracle's restatement of
    StepController.compute_step) registers them here.  ``None, None`` restores the default."""
    global _ERROR_TYPES
    _ERROR_TYPES = None if linear_solver_error is None else (linear_solver_error, step_solver_error)


def _reference_errors():
    """When the reference package is importable, raise ITS exception types so that
    StepController.compute_step (step_control.py:102-104) catches them."""
    if _ERROR_TYPES is not None:
        return _ERROR_TYPES
    try:  # pragma: no cover - depends on the host environment
        from pygradflow.linear_solver import LinearSolverError as RefLSE
        from pygradflow.step.step_solver_error import StepSolverError as RefSSE

        return RefLSE, RefSSE
    except Exception:
        return LinearSolverError, StepSolverError


def _dense(a) -> np.ndarray:
    if hasattr(a, "toarray"):
        a = a.toarray()
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def _dev(a, device) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(device)


_ONE = None


def _one():
    global _ONE
    if _ONE is None:
        _ONE = WorkList.all(1)
    return _ONE


# ------------------------------------------------------------------------------------------------
class B200LinearSolver:
    """LinearSolver(matrix, symmetric=False): the constructor factorises, ``solve`` substitutes.

    method: "lu" (pivoted, default for unsymmetric), "ldlt" (symmetric, inertia available), None =
    LDL' for symmetric input with a pivoted-LU fallback when a pivot breaks down or the factor grows, or the
    reference's iterative solvers "gmres" (gmres_solver.py:7-35, honours ``initial_sol`` and ``trans``) and "minres"
    (minres_solver.py:6-24, symmetric matrices only) -- scipy's iterations, one CTA per system."""

    GROWTH_LIMIT = 1e8

    def __init__(self, matrix, symmetric: bool = False, method: Optional[str] = None, device="cuda"):
        self.symmetric = symmetric
        mat = _dense(matrix)
        assert mat.ndim == 2 and mat.shape[0] == mat.shape[1]
        self.N = N = mat.shape[0]
        self.device = device
        self._neg: Optional[int] = None
        self.method = None
        if N == 0:
            return
        LSE, _ = _reference_errors()
        if not np.all(np.isfinite(mat)):
            raise LSE("matrix has non-finite entries")
        want_ldlt = method == "ldlt" or (method is None and symmetric)
        i32 = dict(dtype=torch.int32, device=device)
        self.Nvec = torch.full((1,), N, **i32)
        self.info = torch.zeros((1,), **i32)
        if method in ("gmres", "minres"):
            # GMRESSolver / MINRESSolver (gmres_solver.py:8-10, minres_solver.py:7-10): keep the matrix, iterate in solve
            assert method == "gmres" or symmetric, "MINRES requires a symmetric matrix"
            self.K = _dev(mat, device).reshape(1, N, N)
            rows = K.krylov_scratch_rows(method == "minres")
            self.scratch = torch.zeros((1, rows, N), dtype=torch.float64, device=device)
            self.iters = torch.zeros((1,), **i32)
            self.method = method
            return
        if want_ldlt:
            ld = max(((N + 63) // 64) * 64, 64)
            Kp = np.eye(ld)
            Kp[:N, :N] = mat
            self.K = _dev(Kp, device).reshape(1, ld, ld)
            self.dvec = torch.zeros((1, ld), dtype=torch.float64, device=device)
            nneg = torch.zeros((1,), **i32)
            K.ldlt_factor(self.K, N, self.Nvec, self.dvec, self.info, nneg, None, _one())
            ok = int(self.info.item()) == 0
            if ok:
                growth = float(torch.tril(self.K[0, :N, :N], -1).abs().max().item()) if N > 1 else 0.0
                ok = np.isfinite(growth) and growth <= self.GROWTH_LIMIT
            if ok:
                self.method = "ldlt"
                self._neg = int(nneg.item())
                return
            if method == "ldlt":
                raise LSE("LDL' factorisation broke down (zero pivot or unbounded growth)")
        self.K = _dev(mat, device).reshape(1, N, N)
        self.piv = torch.zeros((1, N), **i32)
        K.lu_factor(self.K, N, self.Nvec, self.piv, self.info, _one())
        if int(self.info.item()) != 0:  # exactly singular: lu_solver.py:15-17
            raise LSE("LU decomposition failed")
        self.method = "lu"

    def solve(self, rhs, trans: bool = False, initial_sol=None) -> np.ndarray:
        rhs = np.asarray(rhs, dtype=np.float64)
        assert rhs.shape == (self.N,)
        if self.N == 0:
            return np.zeros(0)
        ld = self.K.shape[1]
        r = torch.zeros((1, ld), dtype=torch.float64, device=self.device)
        r[0, : self.N] = torch.from_numpy(rhs).to(self.device)
        if self.method in ("gmres", "minres"):
            x0 = None
            if initial_sol is not None:  # a zero-argument callable (gmres_solver.py:15-16)
                x0 = torch.zeros((1, ld), dtype=torch.float64, device=self.device)
                x0[0, : self.N] = torch.from_numpy(np.asarray(initial_sol(), dtype=np.float64)).to(self.device)
            if self.method == "gmres":
                K.gmres_solve(self.K, self.N, self.Nvec, r, x0, None, trans, self.scratch, self.info, self.iters, _one())
            else:
                K.minres_solve(self.K, self.N, self.Nvec, r, x0, self.scratch, self.info, self.iters, _one())
            code = int(self.info.item())
            if code != 0:
                LSE, _ = _reference_errors()
                raise LSE(f"{self.method.upper()} failed with error code {code}")
        elif self.method == "ldlt":
            K.ldlt_solve(self.K, self.N, self.Nvec, r, _one())
        else:
            K.lu_solve(self.K, self.N, self.Nvec, self.piv, r, trans, _one())
        return r[0, : self.N].cpu().numpy()

    def num_neg_eigvals(self) -> Optional[int]:
        return self._neg

    def rcond(self) -> Optional[float]:
        return None


# ------------------------------------------------------------------------------------------------
class _DeviceIterate:
    """Device copies of what the kernels read from a host ``Iterate`` (x, y, grad f, c, J)."""

    def __init__(self, iterate, device):
        problem = iterate.problem
        n, m = problem.num_vars, problem.num_cons
        self.x = _dev(iterate.x, device).reshape(1, n)
        self.y = _dev(iterate.y, device).reshape(1, m)
        self.grad = _dev(iterate.obj_grad, device).reshape(1, n)
        self.cons = _dev(iterate.cons, device).reshape(1, m) if m > 0 else torch.zeros((1, 0), dtype=torch.float64, device=device)
        self.J = _dev(_dense(iterate.cons_jac), device).reshape(1, m, n) if m > 0 else None


def _cache(iterate, device) -> _DeviceIterate:
    c = getattr(iterate, "_b200_cache", None)
    if c is None:
        c = _DeviceIterate(iterate, device)
        try:
            iterate._b200_cache = c
        except Exception:  # iterates that forbid new attributes are simply re-uploaded
            pass
    return c


class B200StepFunc:
    """StepFunc interface of ScaledImplicitFunc (pygradflow/implicit_func.py:202-294) on the GPU."""

    def __init__(self, problem, orig_iterate, dt: float, device="cuda"):
        self.problem = problem
        self.orig_iterate = orig_iterate
        self.dt = dt
        self.lamb = 1.0 / dt
        self.n = problem.num_vars
        self.m = problem.num_cons
        self.device = device
        f64 = dict(dtype=torch.float64, device=device)
        self.lb_d = _dev(problem.var_lb, device).reshape(1, self.n)
        self.ub_d = _dev(problem.var_ub, device).reshape(1, self.n)
        self.dt_d = torch.full((1,), dt, **f64)
        self.lb = self.lamb * np.asarray(problem.var_lb)
        self.ub = self.lamb * np.asarray(problem.var_ub)
        self._dL = torch.zeros((1, self.n), **f64)
        self._F = torch.zeros((1, self.n + self.m), **f64)
        self._act = torch.zeros((1, self.n), dtype=torch.uint8, device=device)
        self._rho = torch.zeros((1,), **f64)

    def _grad_lag(self, it: _DeviceIterate, rho: float):
        self._rho.fill_(rho)
        K.aug_lag_grad(it.J, it.grad, it.cons if self.m > 0 else None, it.y if self.m > 0 else None, self._rho,
                       self._dL, None, None, _one())
        return self._dL

    def compute_active_set(self, iterate, rho: float, tau=None) -> np.ndarray:
        """implicit_func.py:72-74; tau (ActiveSetType Explicit / Smallest / Largest, newton_control.py:60-88) selects
        the tau-variant of the projected point (:237-244) inside gf_residual_tau."""
        it = _cache(iterate, self.device)
        o = _cache(self.orig_iterate, self.device)
        dL = self._grad_lag(it, rho)
        m = self.m
        tau_d = None
        if tau is not None:
            tau_d = torch.full((1,), float(tau), dtype=torch.float64, device=self.device)
        K.residual(it.x, it.y if m > 0 else None, o.x, o.y if m > 0 else None, dL, it.cons if m > 0 else None,
                   self.lb_d, self.ub_d, self.dt_d, True, 0, self._act, None, None, _one(), tau=tau_d)
        return self._act[0].cpu().numpy().astype(bool)

    def projection_initial(self, iterate, rho: float, tau=None) -> np.ndarray:
        """implicit_func.py:233-246: the point whose box position decides the active set."""
        it = _cache(iterate, self.device)
        o = _cache(self.orig_iterate, self.device)
        dL = self._grad_lag(it, rho)
        lamb = self.lamb
        if tau is not None:
            p = (lamb * (1 - tau * lamb)) * it.x + (tau * lamb * lamb) * o.x - (tau * lamb) * dL
        else:
            p = lamb * o.x - dL
        return p[0].cpu().numpy()

    # host-side helpers of the StepFunc base class (implicit_func.py:21-60,80-99), kept for interface completeness
    def compute_active_set_box(self, x, lb, ub) -> np.ndarray:
        return np.logical_or(x < lb - 1e-8, x > ub + 1e-8)

    def project_box(self, x, lb, ub, active_set) -> np.ndarray:
        p = np.copy(x)
        p[active_set] = np.clip(x[active_set], lb[active_set], ub[active_set])
        return p

    def project(self, x, active_set) -> np.ndarray:
        return self.project_box(x, self.lb, self.ub, active_set)

    def apply_project_deriv(self, mat, active_set):
        """keep_rows(mat, inactive) (util.py:27-55): the rows of the active variables become zero."""
        import scipy.sparse as sps

        keep = sps.diags([np.logical_not(active_set).astype(np.float64)], [0])
        return (keep @ sps.csr_matrix(mat)).tocsr()

    def deriv(self, jac, hess, active_set):
        """ScaledImplicitFunc.deriv (implicit_func.py:254-286): F' = [[lamb I + P_I H, P_I J'], [-J, lamb I]] assembled
        by gf_kkt_assemble_full(GF_FORM_SCALED_DERIV); returned as a scipy csc matrix like the reference."""
        import scipy.sparse as sps

        n, m, dev = self.n, self.m, self.device
        N = n + m
        i32 = dict(dtype=torch.int32, device=dev)
        act = torch.from_numpy(np.asarray(active_set, dtype=np.uint8)).to(dev).reshape(1, n).contiguous()
        perm = torch.zeros((1, n), **i32)
        nI = torch.zeros((1,), **i32)
        K.index_sets(act, m, perm, nI, None, _one())
        Hd = _dev(_dense(hess), dev).reshape(1, n, n)
        Jd = _dev(_dense(jac), dev).reshape(1, m, n) if m > 0 else None
        out = torch.zeros((1, N, N), dtype=torch.float64, device=dev)
        K.kkt_assemble_full(Hd, Jd, perm, nI, act, self.dt_d, self._rho, out, K.FORM_SCALED_DERIV, _one())
        return sps.csc_matrix(out[0].cpu().numpy())

    def deriv_at(self, iterate, rho: float, active_set: Optional[np.ndarray] = None):
        """implicit_func.py:288-294 (what GlobalizedNewtonMethod.step reads, newton.py:262)."""
        if active_set is None:
            active_set = self.compute_active_set(iterate, rho)
        return self.deriv(iterate.aug_lag_deriv_xy(), iterate.aug_lag_deriv_xx(rho), active_set)

    def value_at(self, iterate, rho: float, active_set: Optional[np.ndarray] = None) -> np.ndarray:
        it = _cache(iterate, self.device)
        o = _cache(self.orig_iterate, self.device)
        dL = self._grad_lag(it, rho)
        m = self.m
        mode = 0
        if active_set is not None:
            self._act.copy_(torch.from_numpy(np.asarray(active_set, dtype=np.uint8)).to(self.device).reshape(1, self.n))
            mode = 1
        K.residual(it.x, it.y if m > 0 else None, o.x, o.y if m > 0 else None, dL, it.cons if m > 0 else None,
                   self.lb_d, self.ub_d, self.dt_d, True, mode, self._act, self._F, None, _one())
        return self._F[0].cpu().numpy()

    def active_set_at_point(self, p: np.ndarray) -> np.ndarray:
        return np.logical_or(p < self.lb - 1e-8, p > self.ub + 1e-8)

    def value_device(self, iterate, rho: float, active_d: torch.Tensor) -> torch.Tensor:
        """F(iterate) with a device-resident active set; result stays on the device."""
        it = _cache(iterate, self.device)
        o = _cache(self.orig_iterate, self.device)
        dL = self._grad_lag(it, rho)
        m = self.m
        K.residual(it.x, it.y if m > 0 else None, o.x, o.y if m > 0 else None, dL, it.cons if m > 0 else None,
                   self.lb_d, self.ub_d, self.dt_d, True, 1, active_d, self._F, None, _one())
        return self._F


class StepResult:
    """pygradflow/step/solver/step_solver.py:16-63; clip / dx fix-up / diff come from gf_step_finish."""

    def __init__(self, orig_iterate, dx, dy, active_set, rcond=None, xn=None, diff=None):
        self.orig_iterate = orig_iterate
        self.dx = dx
        self.dy = dy
        self.active_set = active_set
        self.rcond = rcond
        self.xn = xn
        self._diff = diff
        self._iterate = None

    @property
    def iterate(self):
        if self._iterate is None:
            o = self.orig_iterate
            cls = type(o)
            yn = o.y - self.dy
            ev = getattr(o, "eval", None)
            if ev is not None:
                self._iterate = cls(o.problem, o.params, self.xn, yn, ev)
            else:
                self._iterate = cls(o.problem, o.params, self.xn, yn)
        return self._iterate

    @property
    def diff(self) -> float:
        return self._diff


class B200StepSolver:
    """StepSolver for ``Params.step_solver``: the Symmetric formulation solved by the CUDA kernels."""

    def __init__(self, problem, params, orig_iterate, dt: float, rho: float, device="cuda",
                 linear: Optional[LinearSolverType] = None):
        assert dt > 0.0 and rho > 0.0
        self.problem = problem
        self.params = params
        self.orig_iterate = orig_iterate
        self.dt = dt
        self.rho = rho
        self.n = problem.num_vars
        self.m = problem.num_cons
        self.device = device
        self._func = B200StepFunc(problem, orig_iterate, dt, device)
        if linear is None:
            linear = getattr(params, "b200_linear_solver", None)
        if linear is None:
            # the reference's own Params.linear_solver_type selects the iterative solvers (linear_solver/__init__.py:15,
            # 35-39); its direct solvers all map to the factorisation the engine picks
            name = getattr(getattr(params, "linear_solver_type", None), "name", "")
            linear = LinearSolverType[name] if name in ("GMRES", "MINRES") else LinearSolverType.Auto
        self.engine = KKTEngine(1, self.n, self.m, device, linear,
                                inertia_correction=bool(getattr(params, "inertia_correction", False)))
        f64 = dict(dtype=torch.float64, device=device)
        self.rho_d = torch.full((1,), rho, **f64)
        self.H = None
        self.J = None
        self._have_active = False
        self._factored = False
        self._xn = torch.zeros((1, self.n), **f64)
        self._yn = torch.zeros((1, self.m), **f64)
        self._dx = torch.zeros((1, self.n), **f64)
        self._dy = torch.zeros((1, self.m), **f64)
        self._diff = torch.zeros((1,), **f64)
        self._active_host = None
        self.num_factorizations = 0

    @property
    def func(self) -> B200StepFunc:
        return self._func

    @property
    def active_set(self) -> np.ndarray:
        assert self._active_host is not None
        return self._active_host

    @property
    def jac(self):
        assert self.J is not None
        return self.J[0].cpu().numpy()

    @property
    def hess(self):
        assert self.H is not None
        return self.H[0].cpu().numpy()

    def reset_deriv(self) -> None:
        self._factored = False

    def estimate_rcond(self, mat=None, solver=None) -> Optional[float]:
        """StepSolver.estimate_rcond (step_solver.py:100-113) for the current factorisation (Dixon's estimator on the
        device, KKTEngine.estimate_rcond); `mat` / `solver` are implied by the engine's state."""
        if not self._factored:
            return None
        try:
            rc = self.engine.estimate_rcond(self.H, self.J, self._func.dt_d, self.rho_d, _one())
        except NotImplementedError:
            return None
        return float(rc[0].item())

    def update_derivs(self, iterate) -> None:
        """scaled_step_solver.py:76-79: J = aug_lag_deriv_xy, H = aug_lag_deriv_xx(rho=0) (multiplier y)."""
        n, m = self.n, self.m
        self.J = _dev(_dense(iterate.aug_lag_deriv_xy()), self.device).reshape(1, m, n) if m > 0 else None
        self.H = _dev(_dense(iterate.aug_lag_deriv_xx(rho=0.0)), self.device).reshape(1, n, n)
        self.reset_deriv()

    def update_active_set(self, active_set: np.ndarray) -> None:
        """scaled_step_solver.py:81-83."""
        a = np.array(active_set, dtype=bool, copy=True)
        assert a.shape == (self.n,)
        self._active_host = a
        self.engine.active.copy_(torch.from_numpy(a.astype(np.uint8)).to(self.device).reshape(1, self.n))
        self.engine.update_active_set(_one())
        self._have_active = True
        self.reset_deriv()

    def linear_solver(self, mat):
        """StepSolver.linear_solver (step_solver.py:94-98) for callers that hand over an explicit matrix."""
        return B200LinearSolver(mat, symmetric=True, device=self.device)

    def solve(self, iterate) -> StepResult:
        """scaled_step_solver.py:85-107 + symmetric_step_solver.py:96-164 + step_solver.py:16-63."""
        assert self._have_active and self.H is not None
        eng, f = self.engine, self._func
        LSE, SSE = _reference_errors()
        if not self._factored:
            eng.factor(self.H, self.J, f.dt_d, self.rho_d, _one())
            self.num_factorizations += 1
            info = int(eng.info.item())
            if info == -2 and eng.inertia_correction:  # symmetric_step_solver.py:146-153
                raise SSE() from LSE("Invalid matrix inertia")
            if info != 0:
                raise SSE() from LSE("KKT factorisation failed")
            self._factored = True
        F = f.value_device(iterate, self.rho, eng.active)
        it = _cache(iterate, self.device)
        m = self.m
        eng.step(self.H, self.J, it.x, it.y if m > 0 else None, F, f.dt_d, self.rho_d, f.lb_d, f.ub_d, self._xn,
                 self._yn if m > 0 else None, self._diff, _one(), dx=self._dx, dy=self._dy if m > 0 else None)
        out = torch.cat([self._xn[0], self._dx[0], self._dy[0], self._diff]).cpu().numpy()
        if eng.solve_can_fail and int(eng.info.item()) != 0:  # gmres_solver.py:32-33 / minres_solver.py:21-22
            raise SSE() from LSE(f"{eng.linear.name} failed with error code {int(eng.info.item())}")
        n = self.n
        xn, dx, dy, diff = out[:n], out[n : 2 * n], out[2 * n : 2 * n + m], float(out[-1])
        rcond = self.estimate_rcond() if getattr(self.params, "report_rcond", False) else None
        return StepResult(iterate, dx, dy, self._active_host, rcond, xn=xn.copy(), diff=diff)
