"""One batched Newton-KKT step: the unit BASELINE.json's metric counts.

For every instance: evaluate the residual of the implicit-Euler equations and the active set at
(x, y), assemble the active-set-reduced regularised KKT system, factorise it, solve, finish the step
(clip, dual back-transform, step length) and evaluate the residual norm at the new point -- i.e.
``newton_method(problem, params, iterate, dt, rho).step(iterate)`` of the reference
(pygradflow/newton.py:63-89,307-323 with a fresh factorisation) followed by the controller's
``||F(next)||`` (pygradflow/step/distance_ratio_control.py:34), SURVEY.md 8a rows a1-a17.
"""

from __future__ import annotations

from typing import Dict, Optional

import torch

from . import kernels as K
from .engine import KKTEngine
from .kernels import WorkList
from .params import LinearSolverType
from .problem import BatchedProblem

PHASES = ("eval", "residual", "assemble", "factor", "solve", "finish", "eval_next")


class NewtonKKTStepper:
    def __init__(self, problem: BatchedProblem, linear: LinearSolverType = LinearSolverType.Auto):
        self.problem = problem
        p = problem
        B, n, m, dev = p.B, p.n, p.m, p.device
        from .params import StepSolverType

        linear = problem.configure(linear, StepSolverType.Symmetric, False)
        self.engine = KKTEngine(B, n, m, dev, linear, band=problem.kkt_band(), stage=problem.kkt_stage_structure())
        f64 = dict(dtype=torch.float64, device=dev)
        self.grad = torch.zeros((B, n), **f64)
        self.cons = torch.zeros((B, m), **f64)
        self.obj = torch.zeros((B,), **f64)
        self.dL = torch.zeros((B, n), **f64)
        self.F = torch.zeros((B, n + m), **f64)
        self.xn = torch.zeros((B, n), **f64)
        self.yn = torch.zeros((B, m), **f64)
        self.gn = torch.zeros((B, n), **f64)
        self.cn = torch.zeros((B, m), **f64)
        self.on = torch.zeros((B,), **f64)
        self.diff = torch.zeros((B,), **f64)
        self.fnorm = torch.zeros((B,), **f64)
        self.dt = torch.zeros((B,), **f64)
        self.Jbuf = [p.alloc_jac() for _ in range(2)] if (m > 0 and not p.jac_constant) else [None, None]
        self.Hbuf = p.alloc_hess() if not p.hess_constant else None
        self.work = WorkList.all(B)
        self.events: Optional[Dict[str, list]] = None

    # -- optional per-phase CUDA-event timing (used by bench.py for the roofline numbers) -----------
    def enable_timing(self):
        self.events = {ph: [] for ph in PHASES}
        self.engine.ldlt_events = []

    def ldlt_ms(self) -> Optional[float]:
        """Mean milliseconds of the gf_ldlt_factor launches alone (inside the `factor` phase), after a synchronize."""
        ev = getattr(self.engine, "ldlt_events", None)
        return sum(a.elapsed_time(b) for a, b in ev) / len(ev) if ev else None

    def _mark(self):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    def phase_ms(self) -> Dict[str, float]:
        """Mean milliseconds per step of every phase (call after a synchronize)."""
        out = {}
        for ph, pairs in self.events.items():
            out[ph] = sum(a.elapsed_time(b) for a, b in pairs) / max(1, len(pairs))
        return out

    def step(self, x, y, lamb, rho):
        """Returns (xn, yn, diff, fnorm, info): tensors owned by the stepper, valid until the next call."""
        prob, eng, w = self.problem, self.engine, self.work
        m = prob.m
        ym = y if m > 0 else None
        cons = self.cons if m > 0 else None
        timing = self.events is not None
        t = [self._mark()] if timing else None
        K.dt_from_lamb(lamb, self.dt)
        prob.eval(x, self.grad, self.cons, self.obj, w)
        J = prob.jac(x, self.Jbuf[0], w) if m > 0 else None
        prob.aug_lag_grad(J, self.grad, cons, ym, rho, self.dL, None, None, w)
        if timing:
            t.append(self._mark())
        K.residual(x, ym, x, ym, self.dL, cons, prob.var_lb, prob.var_ub, self.dt, True, 0, eng.active, self.F, None, w)
        eng.update_active_set(w)
        H = prob.lag_hess(x, ym, self.Hbuf, w)
        if timing:
            t.append(self._mark())
        eng.assemble(H, J, self.dt, rho, w)
        if timing:
            t.append(self._mark())
        self._factor_only(H, J, rho)
        if timing:
            t.append(self._mark())
        if eng.linear == LinearSolverType.BlockTri:  # right-hand side, substitution and step finish in one engine call
            eng.step(H, J, x, ym, self.F, self.dt, rho, prob.var_lb, prob.var_ub, self.xn, self.yn if m > 0 else None,
                     self.diff, w)
            if timing:
                t.append(self._mark())
        else:
            K.kkt_rhs(H, J, eng.perm, eng.nI, self.F, self.dt, rho, eng.rhs, w)
            eng.solve(eng.rhs, w)
            if timing:
                t.append(self._mark())
            K.step_finish(x, ym, eng.rhs, eng.perm, eng.nI, self.F, self.dt, rho, prob.var_lb, prob.var_ub, self.xn,
                          self.yn if m > 0 else None, None, None, self.diff, w)
        if timing:
            t.append(self._mark())
        prob.eval(self.xn, self.gn, self.cn, self.on, w)
        Jn = prob.jac(self.xn, self.Jbuf[1], w) if m > 0 else None
        prob.aug_lag_grad(Jn, self.gn, self.cn if m > 0 else None, self.yn if m > 0 else None, rho, self.dL, None, None, w)
        K.residual(self.xn, self.yn if m > 0 else None, x, ym, self.dL, self.cn if m > 0 else None, prob.var_lb,
                   prob.var_ub, self.dt, False, 0, None, None, self.fnorm, w)
        if timing:
            t.append(self._mark())
            for i, ph in enumerate(PHASES):
                self.events[ph].append((t[i], t[i + 1]))
        return self.xn, self.yn, self.diff, self.fnorm, eng.info

    def _factor_only(self, H, J, rho):
        """Factorise the assembled K (KKTEngine.factor minus the assembly, so the phases time separately)."""
        self.engine.factor_assembled(H, J, self.dt, rho, self.work)
