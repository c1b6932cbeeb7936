"""Batched problem families: device-side twins of the reference's ``Problem`` callbacks
(pygradflow/problem.py:112-192) for B independent instances.

Every family keeps its data in HBM as contiguous float64 tensors with the batch index outermost and
evaluates through the family's CUDA kernels (no per-instance Python callbacks on the hot loop).
Constraints are equalities c(x) = 0 plus variable bounds -- the form the reference's
``ConstrainedProblem`` produces (pygradflow/cons_problem.py:8-55); it is a no-op for such problems.
"""

from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from . import kernels as K
from .kernels import WorkList


def _dev(a, device):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64).to(device).contiguous()


class BatchedProblem:
    """Base: B instances, n variables, m equality constraints, bounds var_lb/var_ub [B, n]."""

    jac_constant = False   # J does not depend on x (linear constraints)
    hess_constant = False  # Hessian of the Lagrangian does not depend on (x, y)

    def __init__(self, var_lb: torch.Tensor, var_ub: torch.Tensor, num_cons: int):
        assert var_lb.shape == var_ub.shape and var_lb.dim() == 2
        self.var_lb = var_lb.contiguous()
        self.var_ub = var_ub.contiguous()
        self.B, self.n = var_lb.shape
        self.m = int(num_cons)
        self.device = var_lb.device

    # -- evaluation interface ------------------------------------------------------------------
    def eval(self, x, grad, cons, obj, work: WorkList) -> None:
        """obj[B], grad[B,n] = grad f(x), cons[B,m] = c(x)."""
        raise NotImplementedError

    def jac(self, x, out, work: WorkList) -> Optional[torch.Tensor]:
        """J(x) [B,m,n]; may return a persistent tensor instead of filling ``out``."""
        raise NotImplementedError

    def lag_hess(self, x, y, out, work: WorkList) -> torch.Tensor:
        """Hessian of the Lagrangian at (x, y) [B,n,n], symmetric."""
        raise NotImplementedError

    def select(self, idx) -> "BatchedProblem":
        """The sub-batch of instances ``idx`` (used to shard the batch across ranks)."""
        raise NotImplementedError

    def kkt_band(self):
        """Optional: (order, half_bandwidth) of an ordering of the n + m KKT unknowns (variables 0..n-1, constraints
        n..n+m-1) under which the KKT matrix of every instance is banded; None for dense families."""
        return None

    def kkt_stage_structure(self):
        """Optional: (S, nx, nu) when the family currently hands out the compact stage layout of its derivatives
        (block-bidiagonal Jacobian Jc [B, S*nx, nx + w], diagonal Hessian Hd [B, n]; include/gradflow_b200.h)."""
        return None

    def configure(self, linear, formulation, globalized: bool):
        """Called once by the drivers before buffers are allocated: lets a family pick the layout of its derivatives
        for the requested linear solver; returns the LinearSolverType the engine should use."""
        return linear

    def fused_family(self) -> Optional[str]:
        """Name of the family's fused persistent solver (one warp / CTA runs a whole instance), if it has one."""
        return None

    def alloc_jac(self) -> Optional[torch.Tensor]:
        """A buffer for ``jac`` (None for families with a constant Jacobian or without constraints)."""
        if self.m == 0 or self.jac_constant:
            return None
        return torch.zeros((self.B, self.m, self.n), dtype=torch.float64, device=self.device)

    def alloc_hess(self) -> Optional[torch.Tensor]:
        if self.hess_constant:
            return None
        return torch.zeros((self.B, self.n, self.n), dtype=torch.float64, device=self.device)

    def aug_lag_grad(self, J, grad, cons, y, rho, dL, jty, jtc, work: WorkList) -> None:
        """Iterate.aug_lag_deriv_x (iterate.py:91-94): dL = grad + J'(rho c + y), optionally J'y and J'c."""
        K.aug_lag_grad(J, grad, cons, y, rho, dL, jty, jtc, work)


class BatchedQP(BatchedProblem):
    """f = x'Hx/2 + g'x, c = Ax + b (the reference's generic QP, tests/pygradflow/qp.py:4-30)."""

    jac_constant = True
    hess_constant = True

    def __init__(self, H, A, g, b, lb, ub, device="cuda"):
        H, g = _dev(H, device), _dev(g, device)
        lb, ub = _dev(lb, device), _dev(ub, device)
        m = 0 if A is None else A.shape[1]
        super().__init__(lb, ub, m)
        self.H, self.g = H, g
        self.A = _dev(A, device) if m > 0 else None
        self.b = _dev(b, device) if m > 0 else None

    def eval(self, x, grad, cons, obj, work):
        K.qp_eval(self.H, self.A, self.g, self.b, x, grad, cons if self.m > 0 else None, obj, work)

    def jac(self, x, out, work):
        return self.A

    def lag_hess(self, x, y, out, work):
        return self.H

    def select(self, idx):
        sel = lambda t: None if t is None else t[idx].contiguous()
        q = object.__new__(BatchedQP)
        BatchedProblem.__init__(q, sel(self.var_lb), sel(self.var_ub), self.m)
        q.H, q.g, q.A, q.b = sel(self.H), sel(self.g), sel(self.A), sel(self.b)
        return q


class BatchedRosenbrock(BatchedProblem):
    """Chained Rosenbrock with per-instance coefficients a, b [B, n-1]; bounds only (cfg2).

    n = 2, a = 1, b = 100 is the reference's tests/pygradflow/rosenbrock.py."""

    def __init__(self, a, b, lb, ub, device="cuda"):
        lb, ub = _dev(lb, device), _dev(ub, device)
        super().__init__(lb, ub, 0)
        self.a, self.b = _dev(a, device), _dev(b, device)
        self._zeroed = set()

    def fused_family(self):
        return "rosen" if self.n <= 64 else None  # gf_rosen_fused_solve

    def eval(self, x, grad, cons, obj, work):
        K.rosen_eval(self.a, self.b, x, grad, obj, work)

    def jac(self, x, out, work):
        return None

    def lag_hess(self, x, y, out, work):
        if out.data_ptr() not in self._zeroed:  # the kernel writes the three diagonals only
            out.zero_()
            self._zeroed.add(out.data_ptr())
        K.rosen_hess(self.a, self.b, x, out, work)
        return out

    def select(self, idx):
        q = object.__new__(BatchedRosenbrock)
        BatchedProblem.__init__(q, self.var_lb[idx].contiguous(), self.var_ub[idx].contiguous(), 0)
        q.a, q.b = self.a[idx].contiguous(), self.b[idx].contiguous()
        q._zeroed = set()
        return q


class BatchedOCP(BatchedProblem):
    """Discretised nonlinear optimal-control problems (cfg4): stage-interleaved variables
    z = (x_1, u_0, ..., x_S, u_{S-1}), equality dynamics, bounds on the controls only.

    Device twin of the oracle's ``OCP`` class.  Two layouts of the derivatives: dense J [B, m, n] / H [B, n, n] (only
    the non-zero entries are rewritten per evaluation; what the generic dense / banded engines read), and -- chosen by
    ``configure`` for LinearSolverType.Auto / BlockTri with the Symmetric formulation -- the compact stage layout
    Jc [B, S*nx, nx + w] / Hd [B, n] of the stage-structured engine (gf_stage_*), which never streams the zeros."""

    def __init__(self, A, Bm, Q, R, xinit, umax, h, device="cuda"):
        A, Bm, Q, R, xinit = (_dev(t, device) for t in (A, Bm, Q, R, xinit))
        B, S, nx, _ = A.shape
        nu = Bm.shape[3]
        f64 = dict(dtype=torch.float64)
        lb1 = torch.cat([torch.full((nx,), -float("inf"), **f64), torch.full((nu,), -float(umax), **f64)]).repeat(S)
        lb = lb1.to(device).expand(B, -1).contiguous()
        super().__init__(lb, -lb, S * nx)
        self.A, self.Bm, self.Q, self.R, self.xinit = A, Bm, Q, R, xinit
        self.S, self.nx, self.nu, self.h, self.umax = S, nx, nu, float(h), float(umax)
        self._zeroed = set()
        self.compact = False

    def configure(self, linear, formulation, globalized):
        from .params import LinearSolverType, StepSolverType

        ok = (self.nx == 8 and self.nu % 4 == 0 and self.nu <= 16 and formulation == StepSolverType.Symmetric
              and not globalized)
        if linear == LinearSolverType.BlockTri and not ok:
            raise ValueError("LinearSolverType.BlockTri needs nx == 8, nu in {4, 8, 12, 16}, the Symmetric formulation and a "
                             "non-globalized Newton method")
        self.compact = ok and linear in (LinearSolverType.Auto, LinearSolverType.BlockTri)
        return LinearSolverType.BlockTri if self.compact else linear

    def kkt_stage_structure(self):
        return (self.S, self.nx, self.nu) if self.compact else None

    def alloc_jac(self):
        if self.compact:
            return torch.zeros((self.B, self.S, 2 * self.nx + self.nu, self.nx), dtype=torch.float64, device=self.device)
        return super().alloc_jac()

    def alloc_hess(self):
        if self.compact:
            return torch.zeros((self.B, self.n), dtype=torch.float64, device=self.device)
        return super().alloc_hess()

    def aug_lag_grad(self, J, grad, cons, y, rho, dL, jty, jtc, work):
        if self.compact:
            K.stage_aug_lag_grad(self.S, self.nx, self.nu, J, grad, cons, y, rho, dL, jty, jtc, work)
        else:
            super().aug_lag_grad(J, grad, cons, y, rho, dL, jty, jtc, work)

    def _zero_once(self, out):
        if out.data_ptr() not in self._zeroed:  # the kernels write the non-zero pattern only
            out.zero_()
            self._zeroed.add(out.data_ptr())

    def eval(self, x, grad, cons, obj, work):
        K.ocp_eval(self.S, self.nx, self.nu, self.h, self.A, self.Bm, self.Q, self.R, self.xinit, x, grad, cons, obj,
                   work)

    def jac(self, x, out, work):
        if self.compact:
            if out.data_ptr() not in self._zeroed:  # first touch: the constant entries, for every instance of the batch
                K.ocp_jac_banded(self.S, self.nx, self.nu, self.h, self.A, self.Bm, x, out, False, WorkList.all(self.B))
                self._zeroed.add(out.data_ptr())
            else:                                   # afterwards only the entries that depend on x
                K.ocp_jac_banded(self.S, self.nx, self.nu, self.h, self.A, self.Bm, x, out, True, work)
            return out
        self._zero_once(out)
        K.ocp_jac(self.S, self.nx, self.nu, self.h, self.A, self.Bm, x, out, work)
        return out

    def lag_hess(self, x, y, out, work):
        if self.compact:
            K.ocp_hess_diag(self.S, self.nx, self.nu, 0.1 * self.h, self.Q, self.R, x, y, out, work)
            return out
        self._zero_once(out)
        K.ocp_hess(self.S, self.nx, self.nu, 0.1 * self.h, self.Q, self.R, x, y, out, work)
        return out

    def kkt_band(self):
        """Stage-interleaved order y_0, z_0, y_1, z_1, ...: c_j couples z_{j-1} (x_j) and z_j, so the half-bandwidth
        is nx + (nx + nu) - 1."""
        S, nx, nu = self.S, self.nx, self.nu
        w = nx + nu
        order = []
        for j in range(S):
            order += [self.n + j * nx + r for r in range(nx)]
            order += [j * w + c for c in range(w)]
        return order, nx + w - 1

    def select(self, idx):
        q = object.__new__(BatchedOCP)
        BatchedProblem.__init__(q, self.var_lb[idx].contiguous(), self.var_ub[idx].contiguous(), self.m)
        for name in ("A", "Bm", "Q", "R", "xinit"):
            setattr(q, name, getattr(self, name)[idx].contiguous())
        q.S, q.nx, q.nu, q.h, q.umax = self.S, self.nx, self.nu, self.h, self.umax
        q._zeroed = set()
        q.compact = self.compact
        return q


class BatchedDense(BatchedProblem):
    """Problem whose derivatives are supplied as dense device tensors by the caller (the batch = 1
    plug-in path evaluates a Python ``Problem`` on the host and uploads grad / cons / J / H here)."""

    def __init__(self, lb, ub, num_cons, device="cuda"):
        super().__init__(_dev(lb, device), _dev(ub, device), num_cons)

    def eval(self, x, grad, cons, obj, work):
        raise RuntimeError("BatchedDense has no device evaluator; derivatives are uploaded by the caller")
