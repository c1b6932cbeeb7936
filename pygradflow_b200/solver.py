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
import torch

from . import kernels as K
from .engine import KKTEngine
from .kernels import WorkList
from .params import NewtonType, Params, PenaltyUpdate
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
    outer_iterations: int       # lock-step iterations executed by the batch
    newton_steps: int           # instance-level Newton-KKT steps executed (sum over instances)

    @property
    def success(self):
        return self.status == 1


class BatchedSolver:
    """``Solver(problem, params).solve(x0, y0)`` for a whole batch on one GPU."""

    def __init__(self, problem: BatchedProblem, params: Optional[Params] = None, sync_every: int = 1,
                 use_graph: bool = True, graph_steps: int = 8):
        self.problem = problem
        self.params = params if params is not None else Params()
        self.sync_every = max(1, int(sync_every))
        self.use_graph = bool(use_graph)
        self.graph_steps = max(1, int(graph_steps))  # captured units replayed between two reads of the running count
        p = problem
        B, n, m, dev = p.B, p.n, p.m, p.device
        self.engine = KKTEngine(B, n, m, dev, self.params.linear_solver_type, band=problem.kkt_band())
        f64 = dict(dtype=torch.float64, device=dev)
        i32 = dict(dtype=torch.int32, device=dev)

        def vec(k):
            return torch.zeros((B, k), **f64)

        def point():  # (x, y, grad, cons, obj)
            return [vec(n), vec(m), vec(n), vec(m), torch.zeros((B,), **f64)]

        self.cur, self.mid, self.fin = point(), point(), point()
        self.dL0, self.dLm, self.jty, self.jtc = vec(n), vec(n), vec(n), vec(n)
        self.F = vec(n + m)
        self.lamb = torch.zeros((B,), **f64)
        self.rho = torch.zeros((B,), **f64)
        self.dt = torch.zeros((B,), **f64)
        self.err_sum = torch.zeros((B,), **f64)
        self.lamb_next = torch.zeros((B,), **f64)
        self.diff1, self.diff2, self.mid_norm = (torch.zeros((B,), **f64) for _ in range(3))
        self.theta = torch.zeros((B,), **f64)
        self.total_res = torch.zeros((B,), **f64)
        self.status = torch.zeros((B,), **i32)
        self.iters = torch.zeros((B,), **i32)
        self.accepted = torch.zeros((B,), **i32)
        self.phase = torch.zeros((B,), **i32)
        self.run = WorkList(torch.zeros((B,), **i32), torch.zeros((1,), **i32), B)
        self.second = WorkList(torch.zeros((B,), **i32), torch.zeros((1,), **i32), B)
        self.Jbuf = [None, None]
        self.Hbuf = [None, None]
        if m > 0 and not p.jac_constant:
            self.Jbuf = [torch.zeros((B, m, n), **f64) for _ in range(2)]
        if not p.hess_constant:
            self.Hbuf = [torch.zeros((B, n, n), **f64) for _ in range(2)]
        self.newton_step_count = torch.zeros((1,), dtype=torch.int64, device=dev)
        self.globalized = None
        if self.params.newton_type == NewtonType.Globalized:
            from .globalized import GlobalizedStepper

            # newton.py:218-304; the line search reads its state back per trial, so this mode runs eagerly
            self.globalized = GlobalizedStepper(problem, self.engine, self.params.newton_tol)
            self.use_graph = False

    # ------------------------------------------------------------------------------------------
    def _y(self, pt):
        return pt[1] if self.problem.m > 0 else None

    def _cons(self, pt):
        return pt[3] if self.problem.m > 0 else None

    def _aug_grad(self, J, pt, dL, jty, jtc, work):
        K.aug_lag_grad(J, pt[2], self._cons(pt), self._y(pt), self.rho, dL, jty, jtc, work)

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

        return BatchedResult(
            x=x.clone(), y=y.clone(), status=self.status.clone(), iterations=self.iters.clone(),
            accepted_steps=self.accepted.clone(), lamb=self.lamb.clone(), rho=self.rho.clone(),
            total_res=self.total_res.clone(), outer_iterations=int(self.iters.max().item()),
            newton_steps=int(self.newton_step_count.item()),
        )

    def _top(self, wtop: WorkList):
        """Termination test on the current iterates (solver.py:180-205,306-308) and the list of running instances."""
        prm, prob = self.params, self.problem
        m = prob.m
        x, y, grad, cons, obj = self.cur
        self._J0 = prob.jac(x, self.Jbuf[0], wtop) if m > 0 else None
        self._aug_grad(self._J0, self.cur, self.dL0, self.jty if m > 0 else None, self.jtc if m > 0 else None, wtop)
        K.check_terminate(x, grad, self._cons(self.cur), self.jty if m > 0 else None,
                          self.jtc if m > 0 else None, obj, prob.var_lb, prob.var_ub, prm.opt_tol, prm.active_tol,
                          prm.local_infeas_tol, prm.obj_lower_limit, prm.iteration_limit, self.iters,
                          self.status, self.total_res, wtop)
        nw = self.run.nwork
        K.build_worklist(self.status, 0, 0, self.run)
        self.run.nwork = nw

    def _step(self, on_iteration, outer):
        """One outer iteration of every running instance: two Newton steps, step-size control, commit."""
        prm, prob, eng = self.params, self.problem, self.engine
        m = prob.m
        run, second = self.run, self.second
        x, y, grad, cons, obj = self.cur
        J0 = self._J0
        full = prm.newton_type == NewtonType.Full
        active_set_newton = prm.newton_type == NewtonType.ActiveSet
        dual_norm = prm.penalty_update == PenaltyUpdate.DualNorm
        lb, ub = prob.var_lb, prob.var_ub
        K.dt_from_lamb(self.lamb, self.dt)

        # ---- first Newton step from (x^, y^) = current iterate
        K.residual(x, self._y(self.cur), x, self._y(self.cur), self.dL0, self._cons(self.cur), lb, ub, self.dt,
                   True, 0, eng.active, self.F, None, run)
        xm, ym, gm, cm, om = self.mid
        ls_failed = None
        if self.globalized is not None:
            # GlobalizedNewtonMethod.step(orig_iterate): derivatives, active set, Newton direction, Armijo search
            st = self.globalized.step(self.cur, self.dL0, self.cur, self.dt, self.rho, xm, ym if m > 0 else None,
                                      self.diff1, run)
            ls_failed = (st == 2) & (eng.info == 0)  # a failed factorisation is a rejected step, not a failed search
            H0 = None
        else:
            H0 = prob.lag_hess(x, self._y(self.cur), self.Hbuf[0], run)
            eng.update_active_set(run)
            eng.factor(H0, J0, self.dt, self.rho, run)
            eng.step(H0, J0, x, self._y(self.cur), self.F, self.dt, self.rho, lb, ub, xm,
                     ym if m > 0 else None, self.diff1, run)
        prob.eval(xm, gm, cm, om, run)
        Jm = prob.jac(xm, self.Jbuf[1], run) if m > 0 else None
        self._aug_grad(Jm, self.mid, self.dLm, None, None, run)
        # ||F_unscaled(mid)|| with the active set recomputed at mid (distance_ratio_control.py:34)
        K.residual(xm, self._y(self.mid), x, self._y(self.cur), self.dLm, self._cons(self.mid), lb, ub, self.dt,
                   False, 0, None, None, self.mid_norm, run)
        K.dr_first(self.status, eng.info, self.dt, self.mid_norm, self.diff1, prm.newton_tol, prm.lamb_red,
                   prm.lamb_min, self.phase, self.lamb_next)
        if ls_failed is not None:  # the reference raises out of Solver.solve (newton.py:294): the instance stops here
            self._line_search_failed(ls_failed)
        K.build_worklist(self.phase, PHASE_SECOND, PHASE_SECOND, second, parent=run)
        second.nwork = run.nwork

        # ---- second Newton step from mid
        Hs, Js = H0, J0
        if full or active_set_newton:
            # Full: active set + derivatives at mid (newton.py:83-89); ActiveSet: active set at mid,
            # derivatives frozen (newton.py:205-215).  Refactoring with an unchanged active set
            # reproduces the same factor, so it is done unconditionally.
            K.residual(xm, self._y(self.mid), x, self._y(self.cur), self.dLm, self._cons(self.mid), lb, ub,
                       self.dt, True, 0, eng.active, self.F, None, second)
            if full:
                Hs = prob.lag_hess(xm, self._y(self.mid), self.Hbuf[1], second)
                Js = Jm
            eng.update_active_set(second)
            eng.factor(Hs, Js, self.dt, self.rho, second)
            # a failed refactorisation rejects the step like the first one would (step_control.py:102-104)
        elif self.globalized is None:
            K.residual(xm, self._y(self.mid), x, self._y(self.cur), self.dLm, self._cons(self.mid), lb, ub,
                       self.dt, True, 1, eng.active, self.F, None, second)
        xf, yf, gf, cf, of = self.fin
        if self.globalized is not None:
            # second step: derivatives / active set at mid, direction from the residual at the ORIGINAL iterate
            st = self.globalized.step(self.cur, self.dL0, self.mid, self.dt, self.rho, xf, yf if m > 0 else None,
                                      self.diff2, second)
            self._mark_failed_second(eng)
            self._line_search_failed((st == 2) & (self.phase == PHASE_SECOND))
        else:
            eng.step(Hs, Js, xm, self._y(self.mid), self.F, self.dt, self.rho, lb, ub, xf,
                     yf if m > 0 else None, self.diff2, second)
        if full or active_set_newton:
            self._mark_failed_second(eng)
        K.dr_second(self.dt, self.diff1, self.diff2, prm.theta_max, prm.log_theta_ref, prm.K_P, prm.K_I,
                    prm.lamb_min, prm.lamb_inc, self.err_sum, self.phase, self.lamb_next, self.theta)
        prob.eval(xf, gf, cf, of, second)

        if on_iteration is not None:
            on_iteration(outer, self)
        # ---- accept / reject, penalty, counters (solver.py:318-378)
        ph = self.phase
        self.newton_step_count += ((ph >= 2) & (ph <= 4)).sum() + ((ph == 3) | (ph == 4)).sum()
        K.commit(self.phase, self.lamb_next, prm.lamb_max, dual_norm, self.mid, self.fin, self.cur, self.lamb,
                 self.rho, self.iters, self.accepted, self.status)

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
