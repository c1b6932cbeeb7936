"""Batched globalized Newton step with Armijo line search.

Restates ``GlobalizedNewtonMethod.step`` (pygradflow/newton.py:242-304) for B instances, quirks included:
the Newton system is assembled at the CURRENT iterate but its right-hand side is the residual at the ORIGINAL
iterate (:248), the step is applied to the original iterate (:298), the sufficient-decrease test carries a '+'
(:283-287), at most 30 halvings (:273).  The reference raises when the search is exhausted (:294); a batch
records ``GF_STATUS_LINE_SEARCH_FAILED`` for that instance instead.

Kernels: gf_merit_grad (matrix-free F'^T F), gf_ls_trial, gf_armijo_residual (fused residual norm + Armijo
test, warp-shuffle reductions) plus the shared evaluation / KKT kernels.
"""

from __future__ import annotations

import torch

from . import kernels as K
from .engine import KKTEngine
from .kernels import WorkList
from .problem import BatchedProblem

MAX_TRIALS = 30  # newton.py:273


class GlobalizedStepper:
    def __init__(self, problem: BatchedProblem, engine: KKTEngine, newton_tol: float):
        self.problem, self.engine, self.newton_tol = problem, engine, float(newton_tol)
        p = problem
        B, n, m, dev = p.B, p.n, p.m, p.device
        f64 = dict(dtype=torch.float64, device=dev)
        i32 = dict(dtype=torch.int32, device=dev)
        v = lambda k: torch.zeros((B, k), **f64)
        s = lambda: torch.zeros((B,), **f64)
        self.dLc, self.dLt = v(n), v(n)
        self.Fc, self.Fo = v(n + m), v(n + m)
        self.xs, self.ys = v(n), v(m)
        self.dx, self.dy = v(n), v(m)
        self.xt, self.yt, self.gt, self.ct, self.ot = v(n), v(m), v(n), v(m), s()
        self.mult = v(m)
        self.res, self.inner, self.alpha, self.diff = s(), s(), s(), s()
        self.trials = torch.zeros((B,), **i32)
        self.state = torch.zeros((B,), **i32)   # 0 searching, 1 accepted, 2 exhausted, 3 not in this call
        self.search = WorkList(torch.zeros((B,), **i32), torch.zeros((1,), **i32), B)
        self.Jc = torch.zeros((B, m, n), **f64) if (m > 0 and not p.jac_constant) else None
        self.Jt = torch.zeros((B, m, n), **f64) if (m > 0 and not p.jac_constant) else None
        self.Hc = torch.zeros((B, n, n), **f64) if not p.hess_constant else None
        self.Hr = torch.zeros((B, n, n), **f64) if not p.hess_constant else None
        self.total_trials = 0

    def step(self, orig, dL_orig, cur, dt, rho, xn, yn, diff, work: WorkList):
        """orig / cur: (x, y, grad, cons, obj) tuples of the original and the current iterate; dL_orig the
        augmented-Lagrangian gradient at the original iterate.  Writes the next iterate into xn / yn and the
        step length into diff; returns the per-instance line-search state tensor."""
        prob, eng = self.problem, self.engine
        m = prob.m
        lb, ub = prob.var_lb, prob.var_ub
        x0, y0 = orig[0], (orig[1] if m > 0 else None)
        xc, yc, gc, cc = cur[0], (cur[1] if m > 0 else None), cur[2], (cur[3] if m > 0 else None)
        # derivatives + active set at the current iterate (newton.py:237-240), F(current) (:253)
        Jc = prob.jac(xc, self.Jc, work) if m > 0 else None
        K.aug_lag_grad(Jc, gc, cc, yc, rho, self.dLc, None, None, work)
        K.residual(xc, yc, x0, y0, self.dLc, cc, lb, ub, dt, True, 0, eng.active, self.Fc, None, work)
        Hc = prob.lag_hess(xc, yc, self.Hc, work)
        eng.update_active_set(work)
        eng.factor(Hc, Jc, dt, rho, work)
        # Newton direction from the residual at the ORIGINAL iterate with this active set (:248)
        K.residual(x0, y0, x0, y0, dL_orig, orig[3] if m > 0 else None, lb, ub, dt, True, 1, eng.active, self.Fo, None,
                   work)
        eng.step(Hc, Jc, x0, y0, self.Fo, dt, rho, lb, ub, self.xs, self.ys if m > 0 else None, self.diff, work,
                 dx=self.dx, dy=self.dy if m > 0 else None)
        # merit value and directional derivative (:254-271); Hessian at multiplier y + rho c (iterate.py:102-110)
        if prob.hess_constant:
            Hr = Hc
        else:
            if m > 0:
                torch.addcmul(yc, cc, rho[:, None], out=self.mult)
            Hr = prob.lag_hess(xc, self.mult if m > 0 else None, self.Hr, work)
        K.merit_grad(Hr, Jc, self.Fc, eng.active, dt, rho, self.dx, self.dy if m > 0 else None, self.res, self.inner,
                     work)
        # line search state: instances outside `work` are parked (3), converged ones return the full step (:256-257)
        self.state.fill_(3)
        K.ls_begin(self.res, self.newton_tol, self.state, self.alpha, self.trials, work)
        early = self.state == 1
        # Every kernel of a trial takes its instance list from the device, so a trial with nobody searching changes
        # nothing: while the stream is being captured into a CUDA graph all MAX_TRIALS trials are recorded; run eagerly,
        # the loop reads the running count back and stops as soon as the search is over.
        capturing = torch.cuda.is_current_stream_capturing()
        parent = None if work.list is None and work.count_dev is None else work
        for _ in range(MAX_TRIALS):
            K.build_worklist(self.state, 0, 0, self.search, parent=parent)
            self.search.nwork = work.nwork
            if not capturing and int(self.search.count_dev.item()) == 0:
                break
            self.total_trials += 1
            sw = self.search
            K.ls_trial(xc, yc, self.dx, self.dy if m > 0 else None, self.alpha, self.xt, self.yt if m > 0 else None, sw)
            prob.eval(self.xt, self.gt, self.ct, self.ot, sw)
            Jt = prob.jac(self.xt, self.Jt, sw) if m > 0 else None
            K.aug_lag_grad(Jt, self.gt, self.ct if m > 0 else None, self.yt if m > 0 else None, rho, self.dLt, None,
                           None, sw)
            K.armijo_residual(self.xt, self.yt if m > 0 else None, x0, y0, self.dLt, self.ct if m > 0 else None, lb, ub,
                              dt, self.res, self.inner, self.newton_tol, MAX_TRIALS, self.alpha, self.trials,
                              self.state, None, sw)
        # result = StepResult(orig, alpha dx, alpha dy) (:298) -- or the un-searched step for converged instances
        a = self.alpha[:, None]
        sdx = torch.where(a == 1.0, self.dx, a * self.dx)
        xr = x0 - sdx
        low, high = xr < lb, xr > ub
        xr = torch.where(low, lb, torch.where(high, ub, xr))
        sdx = torch.where(low, x0 - lb, torch.where(high, x0 - ub, sdx))
        ss = (sdx * sdx).sum(dim=1)
        if m > 0:
            sdy = torch.where(a == 1.0, self.dy, a * self.dy)
            yr = y0 - sdy
            ss = ss + (sdy * sdy).sum(dim=1)
        searched = (self.state == 1) & ~early
        inwork = self.state != 3
        xn.copy_(torch.where(searched[:, None], xr, torch.where((early & inwork)[:, None], self.xs, xn)))
        if m > 0:
            yn.copy_(torch.where(searched[:, None], yr, torch.where((early & inwork)[:, None], self.ys, yn)))
        diff.copy_(torch.where(searched, ss.sqrt(), torch.where(early & inwork, self.diff, diff)))
        return self.state
