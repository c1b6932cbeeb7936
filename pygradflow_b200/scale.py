"""Batched power-of-two scaling: device twin of the reference's ``Scaling`` / ``ScaledProblem`` / ``create_scaling``
(pygradflow/scale.py:48-280), the first stage of its ``Transformation`` (pygradflow/transform.py:13-104).

Every instance carries integer exponents: ``var_weights`` [B, n], ``cons_weights`` [B, m], ``obj_weight`` [B];
x_scaled = ldexp(x, var_weights), c_scaled = ldexp(c, cons_weights), f_scaled = ldexp(f, obj_weight).  All scaled
quantities are exact in FP64 (``gf_ldexp``), so a solve of the scaled problem differs from the reference's only by
the rounding of the Newton/KKT path itself.  The wrapped family keeps evaluating through its own kernels.
"""

from __future__ import annotations

from typing import Optional

import torch

from . import kernels as K
from .kernels import WorkList
from .params import ScalingType
from .problem import BatchedProblem


def _weights_from_nominal(values: torch.Tensor) -> torch.Tensor:
    """scale.py:75-77: 1 - frexp(v).exponent (frexp(0) = (0, 0), i.e. weight 1)."""
    return (1 - torch.frexp(values.to(torch.float64)).exponent).to(torch.int32)


class BatchedScaling:
    """scale.py:48-150 for a batch."""

    def __init__(self, var_weights, cons_weights, obj_weight=None):
        self.var_weights = var_weights.to(torch.int32).contiguous()
        self.cons_weights = cons_weights.to(torch.int32).contiguous()
        B = self.var_weights.shape[0]
        if obj_weight is None:
            obj_weight = torch.zeros((B,), dtype=torch.int32, device=self.var_weights.device)
        self.obj_weight = torch.as_tensor(obj_weight, device=self.var_weights.device).to(torch.int32).expand(B).contiguous()

    # the exponents of the derived quantities (scale.py:131-149)
    @property
    def dual_weights(self):
        return self.cons_weights - self.obj_weight[:, None]

    @property
    def bound_weights(self):
        return self.var_weights - self.obj_weight[:, None]

    def scale_primal(self, x):
        return torch.ldexp(x, self.var_weights)

    def unscale_primal(self, x):
        return torch.ldexp(x, -self.var_weights)

    def scale_dual(self, y):
        return torch.ldexp(y, -self.dual_weights)

    def unscale_dual(self, y):
        return torch.ldexp(y, self.dual_weights)


class BatchedScaled(BatchedProblem):
    """scale.py:153-231: the callbacks of the wrapped family with power-of-two weights applied."""

    def __init__(self, problem: BatchedProblem, scaling: BatchedScaling):
        self.inner, self.scaling = problem, scaling
        vw = scaling.var_weights
        super().__init__(torch.ldexp(problem.var_lb, vw), torch.ldexp(problem.var_ub, vw), problem.m)
        self.jac_constant = problem.jac_constant
        self.hess_constant = problem.hess_constant
        f64 = dict(dtype=torch.float64, device=self.device)
        B, n, m = self.B, self.n, self.m
        self._xo = torch.zeros((B, n), **f64)
        self._yo = torch.zeros((B, m), **f64)
        self._yw = (scaling.cons_weights - scaling.obj_weight[:, None]).contiguous()  # y_orig = ldexp(y, cw - ow)
        self._Jc: Optional[torch.Tensor] = None
        self._Hc: Optional[torch.Tensor] = None
        self._all = WorkList.all(B)

    def _orig(self, x, work):
        K.ldexp(x, self._xo, work, cw=self.scaling.var_weights, sc=-1)
        return self._xo

    def eval(self, x, grad, cons, obj, work):
        s = self.scaling
        self.inner.eval(self._orig(x, work), grad, cons, obj, work)
        K.ldexp(grad, grad, work, cw=s.var_weights, sc=-1, ow=s.obj_weight, so=1)       # :171-177
        if self.m > 0:
            K.ldexp(cons, cons, work, cw=s.cons_weights, sc=1)                          # :179-185
        K.ldexp(obj.view(-1, 1), obj.view(-1, 1), work, ow=s.obj_weight, so=1)          # :166-169

    def _scale_jac(self, Ji, out, work):
        s = self.scaling
        K.ldexp(Ji, out, work, rw=s.cons_weights, sr=1, cw=s.var_weights, sc=-1)        # :187-205
        return out

    def jac(self, x, out, work):
        if self.m == 0:
            return self.inner.jac(self._orig(x, work), out, work)
        if self.jac_constant:
            if self._Jc is None:
                Ji = self.inner.jac(self._orig(x, self._all), None, self._all)
                self._Jc = self._scale_jac(Ji, torch.empty_like(Ji), self._all)
            return self._Jc
        Ji = self.inner.jac(self._orig(x, work), out, work)
        return self._scale_jac(Ji, out if out is not None else Ji, work)

    def _scale_hess(self, Hi, out, work):
        s = self.scaling
        K.ldexp(Hi, out, work, rw=s.var_weights, sr=-1, cw=s.var_weights, sc=-1, ow=s.obj_weight, so=1)  # :207-231
        return out

    def lag_hess(self, x, y, out, work):
        if self.hess_constant:
            if self._Hc is None:
                Hi = self.inner.lag_hess(self._orig(x, self._all), y, None, self._all)
                self._Hc = self._scale_hess(Hi, torch.empty_like(Hi), self._all)
            return self._Hc
        yo = None
        if self.m > 0:
            K.ldexp(y, self._yo, work, cw=self._yw, sc=1)
            yo = self._yo
        Hi = self.inner.lag_hess(self._orig(x, work), yo, out, work)
        return self._scale_hess(Hi, out if out is not None else Hi, work)

    def kkt_band(self):
        return self.inner.kkt_band()


def _scale_symmetric(A: torch.Tensor) -> torch.Tensor:
    """scale.py:12-45 for a batch of dense symmetric matrices [B, N, N]: integer exponents [B, N].  The reference
    accumulates the column sums into an integer array, i.e. sums floor(|a|); an instance that has converged keeps
    Rsca = 0, so iterating until every instance has converged changes nothing for it."""
    a = A.abs()
    B, N, _ = a.shape
    D = torch.zeros((B, N), dtype=torch.int32, device=A.device)
    for _ in range(100):
        R = torch.floor(a).sum(dim=1)
        R = torch.where(R < 1e-10, torch.ones_like(R), R).sqrt()
        rs = (1 - torch.frexp(R).exponent).to(torch.int32)
        if not bool((rs != 0).any().item()):
            return D
        a = torch.ldexp(a, rs[:, :, None] + rs[:, None, :])
        D += rs
    raise RuntimeError("Equilibration failed to converge")


def create_scaling(problem: BatchedProblem, params, scaling_primal=None, scaling_dual=None) -> Optional[BatchedScaling]:
    """scale.py:234-280.  ``params.scaling``: a BatchedScaling or a (var_weights, cons_weights[, obj_weight]) tuple of
    arrays broadcastable to [B, n] / [B, m] / [B]."""
    st = params.scaling_type
    dev = problem.device
    B, n, m = problem.B, problem.n, problem.m
    if params.scaling is not None:
        assert st == ScalingType.Custom
        sc = params.scaling
        if isinstance(sc, BatchedScaling):
            return sc
        vw = torch.as_tensor(sc[0]).to(dev).expand(B, n)
        cw = torch.as_tensor(sc[1]).to(dev).expand(B, m)
        ow = torch.as_tensor(sc[2]).to(dev) if len(sc) > 2 else None
        return BatchedScaling(vw, cw, ow)
    if st == ScalingType.NoScaling:
        return None
    if st == ScalingType.Custom:
        raise ValueError("Custom scaling requires explicit scaling")
    if scaling_primal is None:
        raise ValueError("Primal point required for scaling computation")
    f64 = dict(dtype=torch.float64, device=dev)
    xs = torch.as_tensor(scaling_primal, dtype=torch.float64).to(dev).expand(B, n).contiguous()
    w = WorkList.all(B)
    grad, cons, obj = torch.zeros((B, n), **f64), torch.zeros((B, m), **f64), torch.zeros((B,), **f64)
    problem.eval(xs, grad, cons, obj, w)
    if st == ScalingType.Nominal:
        return BatchedScaling(_weights_from_nominal(xs), _weights_from_nominal(cons))
    J = problem.jac(xs, torch.zeros((B, m, n), **f64), w) if m > 0 else torch.zeros((B, 0, n), **f64)
    if st == ScalingType.GradJac:                                                        # scale.py:79-106
        vw = -_weights_from_nominal(grad.abs())
        pres = torch.ldexp(J.abs(), (-vw)[:, None, :].expand(B, m, n))
        maxv = torch.trunc(pres.max(dim=2).values) if m > 0 else torch.zeros((B, 0), **f64)  # integer array: truncated
        return BatchedScaling(vw, _weights_from_nominal(maxv))
    if st == ScalingType.KKT:                                                            # scale.py:108-119
        if scaling_dual is None:
            raise ValueError("Dual point required for KKT scaling computation")
        ys = torch.as_tensor(scaling_dual, dtype=torch.float64).to(dev).expand(B, m).contiguous()
        H = problem.lag_hess(xs, ys if m > 0 else None, torch.zeros((B, n, n), **f64), w)
        Kk = torch.zeros((B, n + m, n + m), **f64)
        Kk[:, :n, :n] = H
        if m > 0:
            Kk[:, n:, :n] = J
            Kk[:, :n, n:] = J.transpose(1, 2)
        wts = _scale_symmetric(Kk)
        return BatchedScaling(-wts[:, :n], wts[:, n:])
    raise ValueError(f"Unknown scaling type {st}")
