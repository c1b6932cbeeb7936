"""Batched driver: per-instance state machine of the reference's outer loop around the Newton/KKT path.

Restates, for B independent instances advanced in lock-step with masks / work lists:
  Solver.solve main loop + _check_terminate          pygradflow/solver.py:180-205,305-380
  StepController.compute_step (failure => 2*lambda)   pygradflow/step/step_control.py:67-107
  DistanceRatioController.step + LogController        pygradflow/step/distance_ratio_control.py:18-78,
                                                      pygradflow/controller.py:29-77
  newton_method / Simplified / Full / ActiveSet       pygradflow/newton.py:35-89,181-215,307-323
  DualNormUpdate / ConstantPenalty                    pygradflow/penalty.py:36-74
Every instance keeps its own lambda, rho, PI integral, status and counters, so iteration counts and
accept / reject sequences are those of B scalar reference solves; no instance is ever advanced past
its own termination test.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Optional

import numpy as np
import os

import torch

from . import kernels as K
from .engine import KKTEngine
from .kernels import WorkList
from .params import ActiveSetType, NewtonType, Params, PenaltyUpdate, StepControlType, StepSolverType
from .problem import BatchedProblem

PHASE_SECOND = 1


@dataclass
class BatchedResult:
    x: torch.Tensor
    y: torch.Tensor
    status: torch.Tensor        # int32 [B], SolverStatus values (pygradflow/status.py) + 6/7
    iterations: torch.Tensor    # int32 [B]
    accepted_steps: torch.Tensor
    lamb: torch.Tensor
    rho: torch.Tensor
    total_res: torch.Tensor
    rcond: Optional[torch.Tensor]   # report_rcond: Dixon estimate of the last KKT matrix each instance factorised
    outer_iterations: int       # lock-step iterations executed by the batch
    newton_steps: int           # instance-level Newton-KKT steps executed (sum over instances)

    @property
    def success(self):
        return self.status == 1


class BatchedSolver:
    """``Solver(problem, params).solve(x0, y0)`` for a whole batch on one GPU."""

    def __init__(self, problem: BatchedProblem, params: Optional[Params] = None, use_graph: bool = True,
                 graph_steps: int = 8):
        self.problem = problem
        self.params = params if params is not None else Params()
        self.use_graph = bool(use_graph)
        self.graph_steps = max(1, int(graph_steps))  # captured units replayed between two reads of the running count
        p = problem
        B, n, m, dev = p.B, p.n, p.m, p.device
        linear = problem.configure(self.params.linear_solver_type, self.params.step_solver_type,
                                   self.params.newton_type == NewtonType.Globalized)
        self.engine = KKTEngine(B, n, m, dev, linear, band=problem.kkt_band(),
                                formulation=self.params.step_solver_type,
                                inertia_correction=self.params.inertia_correction,
                                stage=problem.kkt_stage_structure())
        # StandardStepSolver works on the unscaled implicit function (its own active-set test, F and F')
        self._standard = self.params.step_solver_type == StepSolverType.Standard
        self._scaled = not self._standard
        self._Hrho = [None, None]
        f64 = dict(dtype=torch.float64, device=dev)
        i32 = dict(dtype=torch.int32, device=dev)

        def vec(k):
            return torch.zeros((B, k), **f64)

        def point():  # (x, y, grad, cons, obj)
            return [vec(n), vec(m), vec(n), vec(m), torch.zeros((B,), **f64)]

        self.cur, self.mid, self.fin = point(), point(), point()
        self.dL0, self.dLm, self.jty, self.jtc = vec(n), vec(n), vec(n), vec(n)
        self.dLf = vec(n)
        self.F = vec(n + m)
        self.lamb = torch.zeros((B,), **f64)
        self.rho = torch.zeros((B,), **f64)
        self.dt = torch.zeros((B,), **f64)
        self.err_sum = torch.zeros((B,), **f64)
        self.lamb_next = torch.zeros((B,), **f64)
        self.diff1, self.diff2, self.mid_norm = (torch.zeros((B,), **f64) for _ in range(3))
        self.fin_norm, self.orig_norm = (torch.zeros((B,), **f64) for _ in range(2))
        self.loop_key = torch.zeros((B,), **i32)
        self.exact_curr = torch.zeros((B,), **f64)
        self.theta = torch.zeros((B,), **f64)
        self.tau = torch.zeros((B,), **f64)
        self._tau = None
        self.total_res = torch.zeros((B,), **f64)
        self.status = torch.zeros((B,), **i32)
        self.iters = torch.zeros((B,), **i32)
        self.accepted = torch.zeros((B,), **i32)
        self.phase = torch.zeros((B,), **i32)
        self.run = WorkList(torch.zeros((B,), **i32), torch.zeros((1,), **i32), B)
        self.second = WorkList(torch.zeros((B,), **i32), torch.zeros((1,), **i32), B)
        self.Jbuf = [None, None, None]
        self.Hbuf = [None, None]
        exact = self.params.step_control_type == StepControlType.Exact
        pu = self.params.penalty_update
        self._filter_kind = {PenaltyUpdate.ObjectiveFilter: 0, PenaltyUpdate.LagrangianFilter: 1}.get(pu)
        self._pareto = pu == PenaltyUpdate.ParetoDecrease
        if self._filter_kind is not None:
            cap = int(self.params.penalty_filter_capacity)
            self.rho_pen = torch.zeros((B,), **f64)       # the strategy's own rho (penalty.py:177,209)
            self.filt = torch.zeros((B, cap, 2), **f64)   # PenaltyFilter.entries per instance
            self.nfilt = torch.zeros((B,), **i32)
            self.filt_overflow = torch.zeros((B,), **i32)
            self.dLpm, self.dLpf = (vec(n), vec(n)) if self._filter_kind == 1 else (None, None)
        lag_filter = self._filter_kind == 1
        if m > 0 and not p.jac_constant:
            self.Jbuf = [p.alloc_jac() for _ in range(3 if (exact or lag_filter) else 2)]
        if not p.hess_constant:
            self.Hbuf = [p.alloc_hess() for _ in range(2)]
        if self._standard and m > 0:
            self._Hrho = [torch.empty((B, n, n), **f64) for _ in range(2)]
        self.newton_step_count = torch.zeros((1,), dtype=torch.int64, device=dev)
        self.rcond = None
        if self.params.report_rcond:  # reads sizes back to the host: no graph replay in this mode
            self.rcond = torch.full((B,), float("nan"), **f64)
            self.use_graph = False
        self.globalized = None
        if self.params.newton_type == NewtonType.Globalized:
            if self._standard:
                raise ValueError("NewtonType.Globalized is built on the scaled step formulations only")
            from .globalized import GlobalizedStepper

            # newton.py:218-304; under CUDA-graph capture all 30 Armijo trials are recorded (device work lists make the
            # superfluous ones no-ops), run eagerly the search stops at the running count
            self.globalized = GlobalizedStepper(problem, self.engine, self.params.newton_tol)

    # ------------------------------------------------------------------------------------------
    def _y(self, pt):
        return pt[1] if self.problem.m > 0 else None

    def _cons(self, pt):
        return pt[3] if self.problem.m > 0 else None

    def _aug_grad(self, J, pt, dL, jty, jtc, work):
        self.problem.aug_lag_grad(J, pt[2], self._cons(pt), self._y(pt), self.rho, dL, jty, jtc, work)

    def solve(self, x0=None, y0=None, on_iteration: Optional[Callable] = None,
              max_outer: Optional[int] = None) -> BatchedResult:
        prm, prob = self.params, self.problem
        B, n, m = prob.B, prob.n, prob.m
        dev = prob.device
        x, y, grad, cons, obj = self.cur
        # transform.py:29-54: x0 None -> clip(0, lb, ub); scalars broadcast
        if x0 is None:
            x.copy_(torch.minimum(torch.maximum(torch.zeros_like(x), prob.var_lb), prob.var_ub))
        else:
            x.copy_(torch.as_tensor(x0, dtype=torch.float64).to(dev).expand(B, n))
        if m > 0:
            if y0 is None:
                y.zero_()
            else:
                y.copy_(torch.as_tensor(y0, dtype=torch.float64).to(dev).expand(B, m))
        self.lamb.fill_(prm.lamb_init)
        self.rho.fill_(prm.rho)
        self.err_sum.zero_()
        self.status.zero_()
        self.iters.zero_()
        self.accepted.zero_()
        self.newton_step_count.zero_()
        self.phase.zero_()
        if self._fused_eligible() and on_iteration is None and max_outer is None:
            return self._solve_fused()
        if self._filter_kind is not None:
            self.rho_pen.fill_(prm.rho)
            self.nfilt.zero_()
            self.filt_overflow.zero_()
        allw = WorkList.all(B)
        prob.eval(x, grad, cons, obj, allw)

        run, second = self.run, self.second
        run.nwork = second.nwork = B
        # One unit of the loop = one outer iteration of every running instance (`_step`) followed by the termination
        # test of the new iterates (`_top`, solver.py:306-308).  The unit is launch-bound once most instances have
        # finished (cfg2: a few stragglers run for 20 000 more iterations), so after the first eager unit it is
        # replayed from a CUDA graph: every kernel takes its instance list and count from device memory, hence one
        # captured unit is valid for any state, and a unit with nobody running changes nothing.  The captured grid
        # size (an upper bound of the running count) is re-captured when the count has halved.
        use_graph = self.use_graph and on_iteration is None and max_outer is None
        graph, graph_nwork = None, 0
        self._top(allw)
        outer = 0
        while True:
            nrun = int(run.count_dev.item())  # the only host sync of the loop
            if nrun == 0:
                break
            if max_outer is not None and outer >= max_outer:
                break
            if use_graph and outer >= 1:
                bucket = min(B, max(64, 1 << (nrun - 1).bit_length()))
                if graph is None or bucket < graph_nwork:
                    run.nwork = second.nwork = graph_nwork = bucket
                    graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(graph):
                        self._step(None, 0)
                        self._top(run)
                for _ in range(self.graph_steps):
                    graph.replay()
                outer += self.graph_steps
            else:
                run.nwork = second.nwork = nrun
                self._step(on_iteration, outer)
                self._top(run)
                outer += 1

        if self._filter_kind is not None and bool(self.filt_overflow.any().item()):
            raise RuntimeError("penalty filter capacity exceeded: raise Params.penalty_filter_capacity")
        return BatchedResult(
            x=x.clone(), y=y.clone(), status=self.status.clone(), iterations=self.iters.clone(),
            accepted_steps=self.accepted.clone(), lamb=self.lamb.clone(), rho=self.rho.clone(),
            total_res=self.total_res.clone(), rcond=None if self.rcond is None else self.rcond.clone(),
            outer_iterations=int(self.iters.max().item()),
            newton_steps=int(self.newton_step_count.item()),
        )

    # ---- fused persistent path (cfg2) ----------------------------------------------------------
    FUSED_CHUNK = 4096  # outer iterations per instance and launch (bounds the kernel's run time: ~25 ms)

    def _fused_eligible(self) -> bool:
        """The fused kernel implements the default path only: simplified Newton, DistanceRatio, standard active set,
        Symmetric formulation, a penalty strategy that is a no-op without constraints."""
        prm, prob = self.params, self.problem
        from .params import LinearSolverType

        return (prm.fused and prob.fused_family() == "rosen" and prob.m == 0
                and prm.newton_type == NewtonType.Simplified and prm.step_control_type == StepControlType.DistanceRatio
                and prm.active_set_type == ActiveSetType.Standard and prm.active_set_tau is None
                and prm.step_solver_type == StepSolverType.Symmetric
                and prm.linear_solver_type == LinearSolverType.Auto
                and prm.penalty_update in (PenaltyUpdate.Constant, PenaltyUpdate.DualNorm, PenaltyUpdate.DualEquilibration)
                and not prm.inertia_correction and not prm.report_rcond)

    def _solve_fused(self) -> BatchedResult:
        """Solver.solve of every instance inside gf_rosen_fused_solve (one warp per instance, no lock-step); the host
        only relaunches for the instances that have used up FUSED_CHUNK outer iterations."""
        prm, prob = self.params, self.problem
        B = prob.B
        x, y, grad, cons, obj = self.cur
        p12 = (prm.opt_tol, prm.active_tol, prm.obj_lower_limit, prm.newton_tol, prm.lamb_red, prm.lamb_min, prm.lamb_max,
               prm.lamb_inc, prm.theta_max, prm.log_theta_ref, prm.K_P, prm.K_I)
        nsteps = torch.zeros((B,), dtype=torch.int32, device=prob.device)
        run = self.run
        work, fresh, launches = WorkList.all(B), True, 0
        while True:
            K.rosen_fused_solve(prob.a, prob.b, prob.var_lb, prob.var_ub, x, grad, obj, self.lamb, self.err_sum,
                                self.status, self.iters, self.accepted, nsteps, self.total_res, self.engine.active, p12,
                                prm.iteration_limit, self.FUSED_CHUNK, fresh, work)
            launches += 1
            K.build_worklist(self.status, 0, 0, run)
            nrun = int(run.count_dev.item())
            if nrun == 0:
                break
            run.nwork = nrun
            work, fresh = run, False
        self.fused_launches = launches
        return BatchedResult(
            x=x.clone(), y=y.clone(), status=self.status.clone(), iterations=self.iters.clone(),
            accepted_steps=self.accepted.clone(), lamb=self.lamb.clone(), rho=self.rho.clone(),
            total_res=self.total_res.clone(), rcond=None, outer_iterations=int(self.iters.max().item()),
            newton_steps=int(nsteps.sum().item()),
        )

    def _top(self, wtop: WorkList):
        """Termination test on the current iterates (solver.py:180-205,306-308) and the list of running instances."""
        prm, prob = self.params, self.problem
        m = prob.m
        x, y, grad, cons, obj = self.cur
        self._J0 = prob.jac(x, self.Jbuf[0], wtop) if m > 0 else None
        if self._pareto and m > 0:
            # ParetoDecrease.update (penalty.py:136-168) of the instances that just accepted a step, on the iterate they
            # moved to: it needs J'c and J'y (no rho involved), and dL below must see the new rho
            self._aug_grad(self._J0, self.cur, None, self.jty, self.jtc, wtop)
            K.pareto_update(grad, cons, self.jty, self.jtc, self.phase, self.status, prm.opt_tol, prm.local_infeas_tol,
                            self.rho, wtop)
        self._aug_grad(self._J0, self.cur, self.dL0, self.jty if m > 0 else None, self.jtc if m > 0 else None, wtop)
        K.check_terminate(x, grad, self._cons(self.cur), self.jty if m > 0 else None,
                          self.jtc if m > 0 else None, obj, prob.var_lb, prob.var_ub, prm.opt_tol, prm.active_tol,
                          prm.local_infeas_tol, prm.obj_lower_limit, prm.iteration_limit, self.iters,
                          self.status, self.total_res, wtop)
        nw = self.run.nwork
        K.build_worklist(self.status, 0, 0, self.run)
        self.run.nwork = nw

    def _step(self, on_iteration, outer):
        """One outer iteration of every running instance: Newton step(s), step-size control, commit."""
        prm = self.params
        K.dt_from_lamb(self.lamb, self.dt)
        ctl = prm.step_control_type
        if ctl == StepControlType.DistanceRatio:
            self._control_distance_ratio()
        elif ctl == StepControlType.Exact:
            self._control_exact()
        else:
            self._control_single(fixed=ctl == StepControlType.Fixed)
        if on_iteration is not None:
            on_iteration(outer, self)
        # ---- accept / reject, penalty, counters (solver.py:318-378)
        if self._filter_kind is not None:
            self._penalty_filter()
        # ParetoDecrease acts in _top (on the committed iterate), the filters above: gf_commit leaves rho alone for them
        dual_norm = {PenaltyUpdate.Constant: 0, PenaltyUpdate.DualNorm: 1, PenaltyUpdate.DualEquilibration: 2}.get(
            prm.penalty_update, 0)
        K.commit(self.phase, self.lamb_next, prm.lamb_max, dual_norm, self.mid, self.fin, self.cur, self.lamb,
                 self.rho, self.iters, self.accepted, self.status)

    def _penalty_filter(self):
        """PenaltyFilter.update (penalty.py:201-210) on every accepted candidate + the veto of solver.py:357-378."""
        prob = self.problem
        n, m = prob.n, prob.m
        run = self.run
        if self._filter_kind == 1:
            # LagrangianPenaltyFilter.iterate_entry (penalty.py:228-238): aug_lag_deriv_x of the candidate at the
            # STRATEGY's rho -- of both candidate buffers, the kernel picks per instance by phase
            rho_keep, self.rho = self.rho, self.rho_pen
            const_j = m == 0 or prob.jac_constant
            Jm = self._J0 if const_j else self.Jbuf[1]
            if const_j:
                Jf = self._J0
            elif self.params.step_control_type == StepControlType.Exact:
                Jf = self.Jbuf[2]
            else:
                Jf = prob.jac(self.fin[0], self.Jbuf[2], run)
            self._aug_grad(Jm, self.mid, self.dLpm, None, None, run)
            self._aug_grad(Jf, self.fin, self.dLpf, None, None, run)
            self.rho = rho_keep
        K.filter_update(self._filter_kind, n, m, self.phase, self.status, self.mid[4], self._cons(self.mid), self.dLpm if self._filter_kind == 1 else None,
                        self.fin[4], self._cons(self.fin), self.dLpf if self._filter_kind == 1 else None, self.rho,
                        self.rho_pen, self.filt, self.nfilt, self.filt_overflow, run)

    # ---- Newton steps -------------------------------------------------------------------------
    def _first_newton_step(self):
        """newton_method(...) at the current iterate and its first step -> self.mid; evaluates the problem there
        and the unscaled residual norm |F(mid)| (what every Newton-based controller looks at first)."""
        prm, prob, eng = self.params, self.problem, self.engine
        m = prob.m
        run = self.run
        x, y, grad, cons, obj = self.cur
        J0 = self._J0
        lb, ub = prob.var_lb, prob.var_ub
        tau = self._compute_tau()
        K.residual(x, self._y(self.cur), x, self._y(self.cur), self.dL0, self._cons(self.cur), lb, ub, self.dt,
                   self._scaled, 0, eng.active, self.F, None, run, tau=tau)
        xm, ym, gm, cm, om = self.mid
        ls_failed = None
        if self.globalized is not None:
            # GlobalizedNewtonMethod.step(orig_iterate): derivatives, active set, Newton direction, Armijo search
            st = self.globalized.step(self.cur, self.dL0, self.cur, self.dt, self.rho, xm, ym if m > 0 else None,
                                      self.diff1, run)
            ls_failed = (st == 2) & (eng.info == 0)  # a failed factorisation is a rejected step, not a failed search
            self._H0 = None
        else:
            self._H0 = self._hess(self.cur, J0, 0, run)
            eng.update_active_set(run)
            eng.factor(self._H0, J0, self.dt, self.rho, run)
            if self.rcond is not None:  # step_solver.py:100-113, every time a step solver factorises (diagnostic)
                est = eng.estimate_rcond(self._H0, J0, self.dt, self.rho, run)
                self.rcond.copy_(torch.where((self.status == 0) & (eng.info == 0), est, self.rcond))
            eng.step(self._H0, J0, x, self._y(self.cur), self.F, self.dt, self.rho, lb, ub, xm,
                     ym if m > 0 else None, self.diff1, run)
        self._eval_point(self.mid, self.dLm, 1, self.mid_norm, run)
        return ls_failed

    def _hess(self, pt, J, slot, work):
        """Hessian block the step solver factorises at `pt`: aug_lag_deriv_xx(rho = 0) = lag_hess(x, y) for the scaled
        formulations (scaled_step_solver.py:80-83), aug_lag_deriv_xx(rho) = lag_hess(x, y + rho c) + rho J'J for the
        Standard one (standard_step_solver.py:50-53, iterate.py:103-110)."""
        prob = self.problem
        if not self._standard:
            return prob.lag_hess(pt[0], self._y(pt), self.Hbuf[slot], work)
        if prob.m == 0:
            return prob.lag_hess(pt[0], None, self.Hbuf[slot], work)
        ymod = torch.addcmul(pt[1], self.rho[:, None], pt[3])
        H = prob.lag_hess(pt[0], ymod, self.Hbuf[slot], work)
        out = self._Hrho[slot]
        K.hess_rho(H, J, self.rho, out, work)             # rho J'J + H (own DMMA kernel, gf_syrk.cu)
        return out

    def _compute_tau(self):
        """NewtonController.compute_tau / tau_vals (newton_control.py:40-88) at the current iterate, per instance."""
        prm, prob = self.params, self.problem
        t = prm.active_set_type
        self._tau = None
        if t == ActiveSetType.Standard:
            assert prm.active_set_tau is None
            return None
        assert self.globalized is None, "tau-based active sets with the globalized Newton method are not supported"
        if t == ActiveSetType.Explicit:
            assert prm.active_set_tau is not None
            self.tau.fill_(float(prm.active_set_tau))
        else:
            x, g = self.cur[0], self.dL0
            nonzero = g.abs() > 1e-8                                  # np.isclose(g, 0.0): atol 1e-8
            pos, neg = (g > 0.0) & nonzero, (g < 0.0) & nonzero
            tv = torch.where(pos, (x - prob.var_lb) / g, torch.where(neg, (prob.var_ub - x) / -g, torch.full_like(x, -1.0)))
            if t == ActiveSetType.SmallestActiveSet:
                big = torch.where(tv > 0, tv, torch.full_like(tv, float("inf")))
                mn = big.min(dim=1).values
                self.tau.copy_(torch.where(torch.isinf(mn), torch.ones_like(mn), 0.5 * mn))
            else:
                self.tau.copy_(torch.clamp(tv.max(dim=1).values, min=1.0))
        self._tau = self.tau
        return self._tau

    def _eval_point(self, pt, dL, jslot, norm_out, work):
        """Problem callbacks at `pt`, its augmented-Lagrangian gradient and ||F_unscaled(pt)|| w.r.t. the current
        iterate, active set recomputed at pt (distance_ratio_control.py:34 and the other controllers alike)."""
        prob = self.problem
        m = prob.m
        x, y = self.cur[0], self._y(self.cur)
        prob.eval(pt[0], pt[2], pt[3], pt[4], work)
        J = prob.jac(pt[0], self.Jbuf[jslot], work) if m > 0 else None
        self._aug_grad(J, pt, dL, None, None, work)
        K.residual(pt[0], self._y(pt), x, y, dL, self._cons(pt), prob.var_lb, prob.var_ub, self.dt, False, 0, None,
                   None, norm_out, work)
        return J

    def _next_newton_step(self, src, dL_src, J_src, dst, diff, work):
        """A further step of the same NewtonMethod object from `src` (evaluated, with dL_src / J_src) into `dst`."""
        prm, prob, eng = self.params, self.problem, self.engine
        m = prob.m
        x, y = self.cur[0], self._y(self.cur)
        lb, ub = prob.var_lb, prob.var_ub
        full = prm.newton_type == NewtonType.Full
        active_set_newton = prm.newton_type == NewtonType.ActiveSet
        if self.globalized is not None:
            # derivatives / active set at src, direction from the residual at the ORIGINAL iterate
            st = self.globalized.step(self.cur, self.dL0, src, self.dt, self.rho, dst[0], dst[1] if m > 0 else None,
                                      diff, work)
            return st
        Hs, Js = self._H0, self._J0
        if full or active_set_newton:
            # Full: active set + derivatives at src (newton.py:83-89); ActiveSet: active set at src,
            # derivatives frozen (newton.py:205-215).  Refactoring with an unchanged active set
            # reproduces the same factor, so it is done unconditionally.
            K.residual(src[0], self._y(src), x, y, dL_src, self._cons(src), lb, ub, self.dt, self._scaled, 0,
                       eng.active, self.F, None, work, tau=self._tau)
            if full:
                Hs = self._hess(src, J_src, 1, work)
                Js = J_src
            eng.update_active_set(work)
            eng.factor(Hs, Js, self.dt, self.rho, work)
            # a failed refactorisation rejects the step like the first one would (step_control.py:102-104)
        else:
            K.residual(src[0], self._y(src), x, y, dL_src, self._cons(src), lb, ub, self.dt, self._scaled, 1,
                       eng.active, self.F, None, work)
        eng.step(Hs, Js, src[0], self._y(src), self.F, self.dt, self.rho, lb, ub, dst[0],
                 dst[1] if m > 0 else None, diff, work)
        return None

    # ---- controllers (step_control.py:123-150) ------------------------------------------------
    def _control_distance_ratio(self):
        """DistanceRatioController.step (distance_ratio_control.py:18-78)."""
        prm, prob, eng = self.params, self.problem, self.engine
        run, second = self.run, self.second
        ls_failed = self._first_newton_step()
        K.dr_first(self.status, eng.info, self.dt, self.mid_norm, self.diff1, prm.newton_tol, prm.lamb_red,
                   prm.lamb_min, self.phase, self.lamb_next)
        if ls_failed is not None:  # the reference raises out of Solver.solve (newton.py:294): the instance stops here
            self._line_search_failed(ls_failed)
        K.build_worklist(self.phase, PHASE_SECOND, PHASE_SECOND, second, parent=run)
        second.nwork = run.nwork
        # ---- second Newton step from mid
        Jm = self.Jbuf[1] if (prob.m > 0 and not prob.jac_constant) else self._J0
        st = self._next_newton_step(self.mid, self.dLm, Jm, self.fin, self.diff2, second)
        if self.globalized is not None or prm.newton_type in (NewtonType.Full, NewtonType.ActiveSet) or eng.solve_can_fail:
            self._mark_failed_second(eng)
        if st is not None:
            self._line_search_failed((st == 2) & (self.phase == PHASE_SECOND))
        K.dr_second(self.dt, self.diff1, self.diff2, prm.theta_max, prm.log_theta_ref, prm.K_P, prm.K_I,
                    prm.lamb_min, prm.lamb_inc, self.err_sum, self.phase, self.lamb_next, self.theta)
        xf, yf, gf, cf, of = self.fin
        prob.eval(xf, gf, cf, of, second)
        if os.environ.get("GF_EAGER_COUNT") == "1":  # A/B hook: the counter as the eager elementwise launches it used to be
            ph = self.phase
            self.newton_step_count += ((ph >= 2) & (ph <= 4)).sum() + ((ph == 3) | (ph == 4)).sum()
        else:
            K.count_newton_steps(self.phase, self.newton_step_count)  # one launch (was a dozen eager elementwise kernels)

    def _control_single(self, fixed: bool):
        """ResiduumRatioController.step (residuum_ratio_control.py:18-63) / FixedStepSizeController.step
        (fixed_control.py:12-19): one Newton step per outer iteration; the verdict and the log-PI update of lambda run in
        gf_single_control."""
        prm, prob, eng = self.params, self.problem, self.engine
        run = self.run
        ls_failed = self._first_newton_step()
        if not fixed:
            x, y = self.cur[0], self._y(self.cur)
            K.residual(x, y, x, y, self.dL0, self._cons(self.cur), prob.var_lb, prob.var_ub, self.dt, False, 0, None,
                       None, self.orig_norm, run)
        K.single_control(fixed, self.status, eng.info, self.dt, self.mid_norm, None if fixed else self.orig_norm, prm,
                         self.err_sum, self.phase, self.lamb_next, self.theta, self.newton_step_count)
        if ls_failed is not None:
            self._line_search_failed(ls_failed)

    EXACT_MAX_IT = 10      # exact_control.py:11
    EXACT_RATE_BOUND = 0.5

    def _control_exact(self):
        """ExactController.step (exact_control.py:16-66): Newton steps until ||F|| <= newton_tol (accept, lambda / 2)
        or the contraction rate exceeds 1/2 / ten steps are used up (reject, 2 lambda).  The steps alternate
        between the `mid` and `fin` buffers; an instance that stops at an even step is committed from `mid`
        (phase 2), at an odd step from `fin` (phase 3).  The per-instance verdicts run in gf_exact_control; the list
        of instances still inside the loop is rebuilt on the device after every stage."""
        prm, prob, eng = self.params, self.problem, self.engine
        run, loop = self.run, self.second
        x, y = self.cur[0], self._y(self.cur)
        K.residual(x, y, x, y, self.dL0, self._cons(self.cur), prob.var_lb, prob.var_ub, self.dt, False, 0, None, None,
                   self.orig_norm, run)
        ls_failed = self._first_newton_step()
        live, curr = self.loop_key, self.exact_curr
        ctl = lambda mode, i, last, val: K.exact_control(mode, i, last, self.status, eng.info, self.dt, val, self.orig_norm,
                                                         prm.newton_tol, self.EXACT_RATE_BOUND, curr, live, self.phase,
                                                         self.lamb_next, self.newton_step_count)
        ctl(0, 0, False, None)
        if ls_failed is not None:
            self._line_search_failed(ls_failed)
            live.copy_(torch.where(ls_failed, torch.zeros_like(live), live))
        pts = (self.mid, self.fin)
        dLs = (self.dLm, self.dLf)
        norms = (self.mid_norm, self.fin_norm)
        diffs = (self.diff1, self.diff2)
        refactors = (self.globalized is not None or prm.newton_type in (NewtonType.Full, NewtonType.ActiveSet)
                     or eng.solve_can_fail)
        for i in range(self.EXACT_MAX_IT):
            last = i == self.EXACT_MAX_IT - 1
            ctl(1, i, last, norms[i % 2])
            if last:
                break
            K.build_worklist(live, 1, 1, loop, parent=run)
            loop.nwork = run.nwork
            src, dst = pts[i % 2], pts[(i + 1) % 2]
            Jsrc = self.Jbuf[1 + (i % 2)] if (prob.m > 0 and not prob.jac_constant) else self._J0
            st = self._next_newton_step(src, dLs[i % 2], Jsrc, dst, diffs[(i + 1) % 2], loop)
            # a failed refactorisation (Full / ActiveSet / Globalized) or iterative solve ends the loop like a StepSolverError
            if refactors:
                ctl(2, i, False, None)
            if st is not None:
                lsf = (live != 0) & (st == 2)
                self._line_search_failed(lsf)
                live.copy_(torch.where(lsf, torch.zeros_like(live), live))
            self._eval_point(dst, dLs[(i + 1) % 2], 1 + ((i + 1) % 2), norms[(i + 1) % 2], loop)

    def _line_search_failed(self, mask):
        """Armijo search exhausted (newton.py:294 raises a bare Exception that ends Solver.solve): the instance
        terminates with GF_STATUS_LINE_SEARCH_FAILED and is not committed."""
        self.status.copy_(torch.where(mask & (self.status == 0), torch.full_like(self.status, 7), self.status))
        self.phase.copy_(torch.where(mask, torch.zeros_like(self.phase), self.phase))

    def _mark_failed_second(self, eng):
        """Solver failure during the second step's refactorisation: reject, lambda <- 2 lambda."""
        failed = (self.phase == PHASE_SECOND) & (eng.info != 0)
        self.lamb_next.copy_(torch.where(failed, 2.0 * (1.0 / self.dt), self.lamb_next))
        self.phase.copy_(torch.where(failed, torch.full_like(self.phase, 5), self.phase))

    # ------------------------------------------------------------------------------------------
    def bounds_dual(self) -> torch.Tensor:
        """Iterate.bounds_dual (iterate.py:136-149) of the current iterates, for result packaging."""
        prm, prob = self.params, self.problem
        x, y, grad, cons, obj = self.cur
        r = -(grad + (self.jty if prob.m > 0 else 0.0))
        atl = (x - prob.var_lb).abs() <= prm.active_tol
        atu = (prob.var_ub - x).abs() <= prm.active_tol
        both = atl & atu
        d = torch.zeros_like(x)
        d = torch.where(atu & ~both, torch.clamp(r, min=0.0), d)
        d = torch.where(atl & ~both, torch.clamp(r, max=0.0), d)
        d = torch.where(both, r, d)
        return d
