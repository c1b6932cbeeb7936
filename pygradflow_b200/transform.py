"""Batched slack transform: problems with general constraint bounds cl <= c(x) <= cu in front of the Newton/KKT
path, which works on equalities + variable bounds.

Device twin of the reference's ``ConstrainedProblem`` (pygradflow/cons_problem.py:8-173) and of the slack part of
``Transformation`` (pygradflow/transform.py:13-104; the scaling stage is pygradflow_b200/scale.py): one slack variable per inequality
row, c_i(x) - s_i = 0 with cl_i <= s_i <= cu_i; equality rows are shifted by -cl_i.  The wrapped family keeps
evaluating through its own kernels; the wrapper only does the index bookkeeping (strided copies of the family's
outputs into the augmented tensors, the constant -1 columns of the Jacobian, the zero slack block of the Hessian).
The slack pattern (which rows are equalities) must be the same for every instance of the batch.
"""

from __future__ import annotations

from typing import Optional

import torch

from . import kernels as K
from .kernels import WorkList
from .problem import BatchedProblem


class BatchedConstrained(BatchedProblem):
    def __init__(self, problem: BatchedProblem, cons_lb: torch.Tensor, cons_ub: torch.Tensor):
        dev = problem.device
        cl = torch.as_tensor(cons_lb, dtype=torch.float64).to(dev).expand(problem.B, problem.m).contiguous()
        cu = torch.as_tensor(cons_ub, dtype=torch.float64).to(dev).expand(problem.B, problem.m).contiguous()
        eq = cl == cu                                                    # cons_problem.py:38-46
        assert bool((eq == eq[0:1]).all().item()), "the equality / inequality pattern must not depend on the instance"
        self.inner = problem
        self.cons_lb, self.cons_ub = cl, cu
        self.slack_positions = torch.nonzero(~eq[0]).flatten()          # device int64
        self.ns = int(self.slack_positions.numel())
        offs = torch.where(eq & (cl != 0.0), -cl, torch.zeros_like(cl))
        self.cons_offsets = offs if bool((offs != 0.0).any().item()) else None
        sp = self.slack_positions
        super().__init__(torch.cat([problem.var_lb, cl[:, sp]], dim=1), torch.cat([problem.var_ub, cu[:, sp]], dim=1),
                         problem.m)                                      # :14-29
        self.no = problem.n
        self.jac_constant = problem.jac_constant
        self.hess_constant = problem.hess_constant
        f64 = dict(dtype=torch.float64, device=dev)
        B, n, m = self.B, self.no, self.m
        self._xo = torch.zeros((B, n), **f64)
        self._go = torch.zeros((B, n), **f64)
        # Scratch of the augmented outputs: every callback fills its scratch for the whole batch and then copies only
        # the rows of the instances in `work` into the caller's tensor (gf_ldexp with no weights is a work-list copy),
        # so that instances outside the list -- e.g. the ones ExactController already stopped -- keep their values.
        self._ga = torch.zeros((B, self.n), **f64)
        self._ci = torch.zeros((B, m), **f64)
        self._Ji = torch.zeros((B, m, n), **f64) if (m > 0 and not problem.jac_constant) else None
        self._Hi = torch.zeros((B, n, n), **f64) if not problem.hess_constant else None
        self._Ja = torch.zeros((B, m, self.n), **f64) if (m > 0 and not problem.jac_constant) else None
        self._Ha = torch.zeros((B, self.n, self.n), **f64) if not problem.hess_constant else None
        self._Jc: Optional[torch.Tensor] = None   # persistent augmented J / H of constant-derivative families
        self._Hc: Optional[torch.Tensor] = None
        self._prepared = set()
        self._all = WorkList.all(B)

    # -- helpers ---------------------------------------------------------------------------------
    def _orig(self, x):
        self._xo.copy_(x[:, : self.no])
        return self._xo

    def _prepare_jac(self, out):
        if out.data_ptr() not in self._prepared:   # constant part: zero, and -1 in (slack row, slack column)
            out.zero_()
            out[:, self.slack_positions, self.no + torch.arange(self.ns, device=out.device)] = -1.0
            self._prepared.add(out.data_ptr())

    def _prepare_hess(self, out):
        if out.data_ptr() not in self._prepared:   # zero slack block (cons_problem.py:124-128)
            out.zero_()
            self._prepared.add(out.data_ptr())

    # -- Problem callbacks -----------------------------------------------------------------------
    def eval(self, x, grad, cons, obj, work):
        ci = self._ci if self.m > 0 else cons
        self.inner.eval(self._orig(x), self._go, ci, obj, work)         # :62-94
        self._ga[:, : self.no].copy_(self._go)                           # slack part of the scratch stays zero
        K.ldexp(self._ga, grad, work)
        if self.m > 0:
            if self.cons_offsets is not None:
                ci.add_(self.cons_offsets)
            if self.ns > 0:
                ci[:, self.slack_positions] -= x[:, self.no:]
            K.ldexp(ci, cons, work)

    def jac(self, x, out, work):                                         # :96-113
        if self.jac_constant:
            if self._Jc is None:
                Ji = self.inner.jac(self._orig(x), None, self._all)
                self._Jc = torch.zeros((self.B, self.m, self.n), dtype=torch.float64, device=self.device)
                self._prepare_jac(self._Jc)
                self._Jc[:, :, : self.no].copy_(Ji)
            return self._Jc
        self._prepare_jac(self._Ja)
        Ji = self.inner.jac(self._orig(x), self._Ji, work)
        self._Ja[:, :, : self.no].copy_(Ji)
        K.ldexp(self._Ja, out, work)
        return out

    def lag_hess(self, x, y, out, work):                                 # :115-128
        if self.hess_constant:
            if self._Hc is None:
                Hi = self.inner.lag_hess(self._orig(x), y, None, self._all)
                self._Hc = torch.zeros((self.B, self.n, self.n), dtype=torch.float64, device=self.device)
                self._Hc[:, : self.no, : self.no].copy_(Hi)
            return self._Hc
        self._prepare_hess(self._Ha)
        Hi = self.inner.lag_hess(self._orig(x), y, self._Hi, work)
        self._Ha[:, : self.no, : self.no].copy_(Hi)
        K.ldexp(self._Ha, out, work)
        return out

    # -- solution transforms ---------------------------------------------------------------------
    def transform_sol(self, x, y):
        """(x, y) of the original problem -> (x with slacks clipped into their bounds, y)  (:130-157)."""
        B, n, m = self.B, self.no, self.m
        f64 = dict(dtype=torch.float64, device=self.device)
        go, c, o = torch.zeros((B, n), **f64), torch.zeros((B, m), **f64), torch.zeros((B,), **f64)
        self.inner.eval(x.contiguous(), go, c, o, self._all)
        sp = self.slack_positions
        s = torch.minimum(torch.maximum(c[:, sp], self.cons_lb[:, sp]), self.cons_ub[:, sp])
        return torch.cat([x, s], dim=1).contiguous(), y

    def restore_sol(self, x, y):
        return x[:, : self.no].contiguous(), y                          # :159-173


def solve_general(problem: BatchedProblem, cons_lb, cons_ub, params=None, x0=None, y0=None):
    """``Solver(problem, params).solve(x0, y0)`` for a batch whose constraints carry bounds: the reference's
    ``Transformation`` (transform.py:13-104) -- initial point (:29-54: None -> clip(0, lb, ub) / 0, scalars
    broadcast), optional power-of-two scaling (scale.py, ``params.scaling_type``), slack transform, solve, restore."""
    from .params import Params
    from .scale import BatchedScaled, create_scaling
    from .solver import BatchedSolver

    params = params if params is not None else Params()
    dev = problem.device
    B, n, m = problem.B, problem.n, problem.m
    f64 = dict(dtype=torch.float64, device=dev)
    if x0 is None:
        x = torch.minimum(torch.maximum(torch.zeros((B, n), **f64), problem.var_lb), problem.var_ub)
    else:
        x = torch.as_tensor(x0, dtype=torch.float64).to(dev).expand(B, n).contiguous()
    y = torch.zeros((B, m), **f64) if y0 is None else torch.as_tensor(y0, dtype=torch.float64).to(dev).expand(B, m).contiguous()
    cl = torch.as_tensor(cons_lb, dtype=torch.float64).to(dev).expand(B, m)
    cu = torch.as_tensor(cons_ub, dtype=torch.float64).to(dev).expand(B, m)
    scaling = create_scaling(problem, params, params.scaling_primal, params.scaling_dual)
    if scaling is not None:                                             # transform.py:56-64, 76-88
        problem = BatchedScaled(problem, scaling)
        cl, cu = torch.ldexp(cl, scaling.cons_weights), torch.ldexp(cu, scaling.cons_weights)   # scale.py:161-162
        x, y = scaling.scale_primal(x).contiguous(), scaling.scale_dual(y).contiguous()
    cp = BatchedConstrained(problem, cl, cu)
    xt, yt = cp.transform_sol(x, y)
    res = BatchedSolver(cp, params).solve(xt, yt)
    res.x_slack = res.x
    res.x, res.y = cp.restore_sol(res.x, res.y)
    if scaling is not None:                                             # transform.py:90-104
        res.x, res.y = scaling.unscale_primal(res.x), scaling.unscale_dual(res.y)
    res.scaling = scaling
    return res
