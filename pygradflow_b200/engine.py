"""Batched KKT step engine: the device twin of the reference's ``SymmetricStepSolver`` +
``LinearSolver`` pair for B instances at once.

Reference call order it reproduces (per instance):
  StepSolver.update_active_set / update_derivs   scaled_step_solver.py:76-83
  ScaledStepSolver.solve -> initial_rhs -> solve_scaled -> _solve_active_set -> _compute_deriv
      -> linear_solver(K) -> LinearSolver.solve(rhs)              symmetric_step_solver.py:96-164
  StepResult (clip, diff)                                           step_solver.py:16-63

All buffers are preallocated once per (B, n, m) and reused across iterations (SURVEY 8b "Ownership").
"""

from __future__ import annotations

from typing import Optional

import torch

from . import kernels as K
from .kernels import WorkList
from .params import LinearSolverType, StepSolverType

SMALL_N = 112  # order up to which the shared-memory-resident pivoted LU is used under Auto


def _roundup(a: int, b: int) -> int:
    return ((a + b - 1) // b) * b


BAND_MIN_N = 256  # order from which Auto prefers the family's banded ordering over the dense LDL'


class KKTEngine:
    def __init__(self, B: int, n: int, m: int, device, linear: LinearSolverType = LinearSolverType.Auto, band=None,
                 formulation: StepSolverType = StepSolverType.Symmetric, inertia_correction: bool = False, stage=None):
        """band: optional (order, half_bandwidth) of the family's KKT ordering (BatchedProblem.kkt_band).
        formulation: Symmetric = the reduced KKT system; Asymmetric / Extended = the full-order unsymmetric systems
        of asymmetric_step_solver.py / extended_step_solver.py, Standard = the derivative of the unscaled implicit
        function (standard_step_solver.py; the caller hands in H_rho and the unscaled residual) -- always pivoted LU,
        order n + m for every instance.
        inertia_correction: Params.inertia_correction (symmetric_step_solver.py:146-153) -- a factorisation whose number
        of negative pivots differs from m fails like a LinearSolverError (the step is rejected, lambda doubled); needs a
        factorisation that reports the inertia (LDL' / Banded), there is no pivoted-LU fallback in this mode."""
        self.B, self.n, self.m = B, n, m
        self.inertia_correction = bool(inertia_correction)
        self.device = device
        N = n + m
        self.form = {StepSolverType.Symmetric: K.FORM_SYMMETRIC, StepSolverType.Asymmetric: K.FORM_ASYMMETRIC,
                     StepSolverType.Extended: K.FORM_EXTENDED, StepSolverType.Standard: K.FORM_STANDARD}[formulation]
        self.iterative = linear in (LinearSolverType.GMRES, LinearSolverType.MINRES)
        if self.form != K.FORM_SYMMETRIC:
            if linear not in (LinearSolverType.Auto, LinearSolverType.LU, LinearSolverType.GMRES):
                # minres_solver.py:9 asserts a symmetric matrix; the symmetric factorisations do not apply either
                raise ValueError(f"step_solver_type={formulation.name} has an unsymmetric matrix: LU or GMRES only")
            if linear != LinearSolverType.GMRES:
                linear = LinearSolverType.LU
        if linear == LinearSolverType.BlockTri:
            self.linear = linear
            self._init_blocktri(stage)
            return
        if linear == LinearSolverType.Auto:
            if band is not None and N >= BAND_MIN_N:
                linear = LinearSolverType.Banded
            elif self.inertia_correction:
                linear = LinearSolverType.LDLT
            else:
                linear = LinearSolverType.LU if N <= SMALL_N else LinearSolverType.LDLT
        if self.inertia_correction and (linear == LinearSolverType.LU or self.iterative):
            # LUSolver.num_neg_eigvals() is None: symmetric_step_solver.py:149-150 raises the same way
            raise Exception("Inertia correction requested but not available")
        self.linear = linear
        if linear == LinearSolverType.Banded:
            self._init_banded(band)
            return
        self.ld = max(_roundup(N, 64), 64) if linear == LinearSolverType.LDLT else max(N, 1)
        f64 = dict(dtype=torch.float64, device=device)
        i32 = dict(dtype=torch.int32, device=device)
        self.K = torch.empty((B, self.ld, self.ld), **f64)
        self.rhs = torch.zeros((B, self.ld), **f64)
        self.perm = torch.zeros((B, n), **i32)
        self.nI = torch.zeros((B,), **i32)
        self.Nvec = torch.zeros((B,), **i32)
        self.piv = torch.zeros((B, self.ld), **i32)
        self.info = torch.zeros((B,), **i32)
        self.active = torch.zeros((B, n), dtype=torch.uint8, device=device)
        if self.form != K.FORM_SYMMETRIC:
            self.perm_id = torch.arange(n, **i32).repeat(B, 1).contiguous()
            self.n_full = torch.full((B,), n, **i32)
            self.N_full = torch.full((B,), N, **i32)
            self.zero_rho = torch.zeros((B,), **f64)
        # Optional (LDL'): gather K inside the factorisation kernels instead of writing it first (gf_kkt_ldlt_factor).
        # Bit-identical, but the dependent index -> H loads in every tile prologue cost the factorisation 2.9 ms while
        # the saved assembly is 3.3 ms (cfg3): the step gains 1 %, the DMMA kernels lose 10 % -- off by default.
        self.fuse_assembly = False
        if linear == LinearSolverType.LDLT:
            self.dvec = torch.zeros((B, self.ld), **f64)
            self.nneg = torch.zeros((B,), **i32)
            self.info_lu = torch.zeros((B,), **i32)
            self.fbkey = torch.zeros((B,), **i32)  # 0: LDL' factor valid, != 0: pivoted-LU fallback
            self._ok = WorkList(torch.zeros((B,), **i32), torch.zeros((1,), **i32), B)
            self._fb = WorkList(torch.zeros((B,), **i32), torch.zeros((1,), **i32), B)
        if self.iterative:
            # gmres_solver.py / minres_solver.py: the "factorisation" only keeps the matrix; every solve runs the whole
            # iteration (one CTA per instance) and may fail (info != 0 -> LinearSolverError -> the step is rejected)
            rows = K.krylov_scratch_rows(linear == LinearSolverType.MINRES)
            self.kscratch = torch.zeros((B, rows, self.ld), **f64)
            self.kiters = torch.zeros((B,), **i32)
        self.n_factor_calls = 0

    @property
    def solve_can_fail(self) -> bool:
        """True when LinearSolver.solve itself can raise (the iterative solvers), not only the constructor."""
        return self.iterative

    def _init_banded(self, band):
        assert band is not None, "LinearSolverType.Banded needs a problem family with kkt_band()"
        B, n, m, device = self.B, self.n, self.m, self.device
        N = n + m
        order, bw = band
        bw = int(bw) | 1  # bw + 1 even: rows move as 16-byte pieces
        assert bw + 1 <= 64 and len(order) == N
        f64 = dict(dtype=torch.float64, device=device)
        i32 = dict(dtype=torch.int32, device=device)
        self.bw = bw
        self.order = torch.as_tensor(order, dtype=torch.int32).to(device).contiguous()
        self.pos = torch.empty_like(self.order)
        self.pos[self.order.long()] = torch.arange(N, **i32)
        self.ld = N
        self.K = None  # no dense KKT matrix in this mode
        self.Kband = torch.empty((B, N, bw + 1), **f64)
        self.bandv = torch.zeros((B, N), **f64)
        self.rhs = torch.zeros((B, N), **f64)
        self.perm = torch.zeros((B, n), **i32)
        self.nI = torch.zeros((B,), **i32)
        self.Nvec = torch.zeros((B,), **i32)
        self.info = torch.zeros((B,), **i32)
        self.nneg = torch.zeros((B,), **i32)
        self.active = torch.zeros((B, n), dtype=torch.uint8, device=device)
        self.n_factor_calls = 0

    def _init_blocktri(self, stage):
        """Stage-structured engine (gf_stage_kkt_factor / gf_stage_kkt_solve): H and J arrive in the family's compact
        layout (BatchedProblem.kkt_stage_structure); no KKT matrix is ever assembled."""
        assert stage is not None, "LinearSolverType.BlockTri needs a family with kkt_stage_structure()"
        B, n, m, device = self.B, self.n, self.m, self.device
        S, nx, nu = stage
        assert n == S * (nx + nu) and m == S * nx
        self.stage = (int(S), int(nx), int(nu))
        f64 = dict(dtype=torch.float64, device=device)
        i32 = dict(dtype=torch.int32, device=device)
        self.ld = n + m
        self.K = None
        self.Tinv, self.Pf, self.Qf = (torch.zeros((B, S, nx * nx), **f64) for _ in range(3))
        self.rhs = torch.zeros((B, n + m), **f64)
        self.perm = torch.zeros((B, n), **i32)
        self.nI = torch.zeros((B,), **i32)
        self.Nvec = torch.zeros((B,), **i32)
        self.info = torch.zeros((B,), **i32)
        self.nneg = torch.zeros((B,), **i32)
        self.active = torch.zeros((B, n), dtype=torch.uint8, device=device)
        self.perm_id = torch.arange(n, **i32).repeat(B, 1).contiguous()
        self.n_full = torch.full((B,), n, **i32)
        self.n_factor_calls = 0

    # ---------------------------------------------------------------------------------------
    def update_active_set(self, work: WorkList, active: Optional[torch.Tensor] = None):
        """np.where(~A) / np.where(A) for the stored active set (self.active unless given)."""
        K.index_sets(self.active if active is None else active, self.m, self.perm, self.nI, self.Nvec, work)

    def assemble(self, H, J, dt, rho, work: WorkList):
        """The reduced symmetric KKT matrix of symmetric_step_solver.py:49-77 in the layout of the factorisation."""
        if self.linear == LinearSolverType.BlockTri:
            return  # the factorisation forms the Schur complement straight from the compact H, J
        if self.form != K.FORM_SYMMETRIC:
            K.kkt_assemble_full(H, J, self.perm, self.nI, self.active, dt, rho, self.K, self.form, work)
        elif self.linear == LinearSolverType.Banded:
            K.band_assemble(H, J, self.active, self.order, self.bw, dt, rho, self.Kband, work)
        elif self.linear == LinearSolverType.LU or self.iterative:
            K.kkt_assemble(H, J, self.perm, self.nI, dt, rho, self.K, 1, False, work)
        elif not self.fuse_assembly:
            K.kkt_assemble(H, J, self.perm, self.nI, dt, rho, self.K, 64, True, work)

    def factor(self, H, J, dt, rho, work: WorkList):
        """Assemble the reduced symmetric KKT matrix and factorise it; per-instance result in self.info."""
        self.assemble(H, J, dt, rho, work)
        self.factor_assembled(H, J, dt, rho, work)

    def factor_assembled(self, H, J, dt, rho, work: WorkList):
        self.n_factor_calls += 1
        Nmax = self.n + self.m
        if self.linear == LinearSolverType.BlockTri:
            S, nx, nu = self.stage
            # info = -2: some H_ii + lamb <= 0, K is not quasi-definite -- reported like a failed factorisation (the step
            # is rejected and lambda doubled, step_control.py:102-104), as in Banded mode
            K.stage_kkt_factor(S, nx, nu, J, H, self.active, dt, rho, self.Tinv, self.Pf, self.Qf, self.info, self.nneg,
                               work)
            return
        if self.linear == LinearSolverType.Banded:
            K.band_factor(self.Kband, self.bw, self.info, self.nneg, work)
            # inertia of a quasi-definite K: exactly m negative pivots (the active rows are +1); anything else is
            # reported like a failed factorisation (the step is rejected and lambda doubled, step_control.py:102-104)
            torch.where((self.info == 0) & (self.nneg != self.m), torch.full_like(self.info, -2), self.info,
                        out=self.info)
            return
        if self.iterative:
            self.info.zero_()  # GMRESSolver / MINRESSolver constructors cannot fail
            return
        if self.linear == LinearSolverType.LU:
            K.lu_factor(self.K, Nmax, self._order(), self.piv, self.info, work)
            return
        ev = getattr(self, "ldlt_events", None)  # optional CUDA-event pairs around the factorisation launches alone
        if ev is not None:
            e0 = torch.cuda.Event(enable_timing=True)
            e0.record()
        if self.fuse_assembly:
            K.kkt_ldlt_factor(H, J, self.perm, self.nI, dt, rho, self.Nvec, self.K, self.dvec, self.info, self.nneg, work)
        else:
            K.ldlt_factor(self.K, Nmax, self.Nvec, self.dvec, self.info, self.nneg, self.nI, work)
        if ev is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            ev.append((e0, e1))
        # Instances whose pivots are not those of a quasi-definite matrix (or hit a zero pivot) are
        # re-assembled in full and factorised with partial pivoting, like the reference's LU.
        self.fbkey.copy_(self.info)
        parent = None if work.list is None and work.count_dev is None else work
        K.build_worklist(self.fbkey, 0, 0, self._fb, parent=parent, invert=True)
        self._fb.nwork = work.nwork
        K.kkt_assemble(H, J, self.perm, self.nI, dt, rho, self.K, 1, False, self._fb)
        K.lu_factor(self.K, Nmax, self.Nvec, self.piv, self.info_lu, self._fb)
        torch.where(self.fbkey != 0, self.info_lu, self.info, out=self.info)
        if self.inertia_correction:
            # symmetric_step_solver.py:146-153: num_neg_eigvals != m => LinearSolverError("Invalid matrix inertia").  By
            # Sylvester's law the signs of D are the inertia whatever their positions, so the unpivoted LDL' counts it even
            # when the pivot pattern is not the quasi-definite one; the SOLUTION of such an instance still comes from the
            # pivoted LU above (stable), only the count from D.  A broken-down LDL' (zero pivot) has no inertia: rejected.
            bad = (self.nneg != self.m) | (self.fbkey > 0) | (self.fbkey == -1)
            torch.where((self.info == 0) & bad, torch.full_like(self.info, -2), self.info, out=self.info)

    def solve(self, rhs, work: WorkList, trans: bool = False):
        """rhs[B, ld] <- K^{-1} rhs for the instances in ``work``."""
        Nmax = self.n + self.m
        if self.linear == LinearSolverType.BlockTri:
            raise NotImplementedError("the stage-structured engine solves from the residual (KKTEngine.step)")
        if self.linear == LinearSolverType.Banded:
            K.band_permute(self.perm, self.nI, self.pos, rhs, self.bandv, True, self.m, work)
            K.band_solve(self.Kband, self.bw, self.bandv, work)  # symmetric: trans is irrelevant
            K.band_permute(self.perm, self.nI, self.pos, rhs, self.bandv, False, self.m, work)
            return
        if self.linear == LinearSolverType.GMRES:
            # AsymmetricStepSolver hands GMRES its start vector (asymmetric_step_solver.py:125-138,154): b0 in the rows
            # of the active variables, which is where rhs already holds it (:106-123)
            mask = self.active if (self.form == K.FORM_ASYMMETRIC and rhs is self.rhs) else None
            K.gmres_solve(self.K, Nmax, self._order(), rhs, None, mask, trans, self.kscratch, self.info, self.kiters, work)
            return
        if self.linear == LinearSolverType.MINRES:
            K.minres_solve(self.K, Nmax, self._order(), rhs, None, self.kscratch, self.info, self.kiters, work)
            return
        if self.linear == LinearSolverType.LU:
            K.lu_solve(self.K, Nmax, self._order(), self.piv, rhs, trans, work)
            return
        parent = None if work.list is None and work.count_dev is None else work
        K.build_worklist(self.fbkey, 0, 0, self._ok, parent=parent, invert=False)
        K.build_worklist(self.fbkey, 0, 0, self._fb, parent=parent, invert=True)
        self._ok.nwork = work.nwork
        self._fb.nwork = work.nwork
        K.ldlt_solve(self.K, Nmax, self.Nvec, rhs, self._ok)  # symmetric: trans is irrelevant
        K.lu_solve(self.K, Nmax, self.Nvec, self.piv, rhs, trans, self._fb)

    def estimate_rcond(self, H, J, dt, rho, work: Optional[WorkList] = None) -> torch.Tensor:
        """``StepSolver.estimate_rcond`` (step_solver.py:100-113) = Dixon's estimator (step/cond_estimate.py:13-114) for
        every instance of the batch, on the CURRENT factorisation: power iterations on A'A (two batched matrix-vector
        products with a re-assembled copy of the matrix) and on its inverse (``solve(trans=True)`` then ``solve``),
        start vectors from ``default_rng(42)`` exactly as the reference draws them.  A diagnostic (``report_rcond``):
        it keeps a second matrix buffer and reads the orders back to the host."""
        if self.linear in (LinearSolverType.Banded, LinearSolverType.BlockTri):
            raise NotImplementedError("report_rcond is not available with the banded / stage-structured factorisations")
        import numpy as np

        B, ld, dev = self.B, self.ld, self.device
        work = work if work is not None else WorkList.all(B)
        f64 = dict(dtype=torch.float64, device=dev)
        if getattr(self, "Kmat", None) is None:
            self.Kmat = torch.zeros((B, ld, ld), **f64)
        else:
            self.Kmat.zero_()
        if self.form != K.FORM_SYMMETRIC:
            K.kkt_assemble_full(H, J, self.perm, self.nI, self.active, dt, rho, self.Kmat, self.form, work)
        else:
            K.kkt_assemble(H, J, self.perm, self.nI, dt, rho, self.Kmat, 1, False, work)
        Nb = self._order().to(torch.int64)                                  # [B]
        Nh = Nb.cpu().numpy().astype(np.float64)
        Nmax = int(Nh.max()) if B > 0 else 0
        col = torch.arange(ld, device=dev)[None, :]
        valid = col < Nb[:, None]
        z = torch.as_tensor(np.random.default_rng(seed=42).normal(size=2 * max(Nmax, 1)), **f64)
        zero = torch.zeros((), **f64)
        x = torch.where(valid, z[col.clamp(max=z.numel() - 1)].expand(B, ld), zero)
        y = torch.where(valid, z[(col + Nb[:, None]).clamp(max=z.numel() - 1)], zero)
        x = x / torch.linalg.vector_norm(x, dim=1, keepdim=True).clamp_min(1e-300)
        y = y / torch.linalg.vector_norm(y, dim=1, keepdim=True).clamp_min(1e-300)
        with np.errstate(divide="ignore"):
            its_h = -2 * np.ceil(np.log((1.0 - 0.99) / 1.6 * np.power(np.maximum(Nh, 1.0), -0.5)) / np.log(10.0))
        its = torch.as_tensor(its_h, **f64)
        xprod, yprod = x.clone(), y.clone()
        xfac, yfac = torch.ones((B,), **f64), torch.ones((B,), **f64)
        Kt = self.Kmat.transpose(1, 2)
        for k in range(int(its_h.max()) if B > 0 else 0):
            live = (its > k)[:, None]
            xn = torch.bmm(Kt, torch.bmm(self.Kmat, xprod.unsqueeze(2))).squeeze(2)
            xn = torch.where(valid, xn, zero)
            buf = yprod.clone()
            self.solve(buf, work, trans=True)
            self.solve(buf, work, trans=False)
            yn = torch.where(valid, buf, zero)
            xnorm = torch.linalg.vector_norm(xn, dim=1)
            ynorm = torch.linalg.vector_norm(yn, dim=1)
            xfac = torch.where(live[:, 0], xfac * xnorm, xfac)
            yfac = torch.where(live[:, 0], yfac * ynorm, yfac)
            xprod = torch.where(live, xn / xnorm[:, None], xprod)
            yprod = torch.where(live, yn / ynorm[:, None], yprod)
        pw = 1.0 / (2.0 * its.clamp_min(1.0))
        xdot = torch.pow((x * xprod).sum(dim=1) * xfac, pw)
        ydot = torch.pow((y * yprod).sum(dim=1) * yfac, pw)
        cond = xdot * ydot
        rc = torch.where(torch.isinf(xdot) | torch.isinf(ydot) | torch.isinf(cond), torch.zeros_like(cond), 1.0 / cond)
        return rc

    def _order(self):
        """Per-instance order of the system handed to the LU: |I| + m (reduced) or n + m (full-order formulations)."""
        return self.Nvec if self.form == K.FORM_SYMMETRIC else self.N_full

    def step(self, H, J, xbase, ybase, F, dt, rho, lb, ub, xn, yn, diff, work: WorkList, dx=None, dy=None):
        """ScaledStepSolver.solve + StepResult for the current factor: rhs, substitution, step finish."""
        if self.linear == LinearSolverType.BlockTri:
            S, nx, nu = self.stage
            K.stage_kkt_solve(S, nx, nu, J, H, self.active, F, dt, rho, self.Tinv, self.Pf, self.Qf, self.rhs, work)
            K.step_finish(xbase, ybase, self.rhs, self.perm_id, self.n_full, F, dt, rho, lb, ub, xn, yn, dx, dy, diff,
                          work)
            return
        if self.form != K.FORM_SYMMETRIC:
            # asymmetric_step_solver.py:140-173 / extended_step_solver.py:85-112: the solution is (dx, sy) itself
            K.kkt_rhs_full(self.n, self.m, self.perm, self.nI, self.active, F, dt, rho, self.rhs, self.form, work)
            self.solve(self.rhs, work)
            # Standard: the solution is (dx, dy) itself (standard_step_solver.py:78-79): rho = 0 makes the dy formula
            # of gf_step_finish, fact (sy - rho F_y), the identity exactly
            K.step_finish(xbase, ybase, self.rhs, self.perm_id, self.n_full, F, dt,
                          self.zero_rho if self.form == K.FORM_STANDARD else rho, lb, ub, xn, yn, dx, dy, diff, work)
            return
        K.kkt_rhs(H, J, self.perm, self.nI, F, dt, rho, self.rhs, work)
        self.solve(self.rhs, work)
        K.step_finish(xbase, ybase, self.rhs, self.perm, self.nI, F, dt, rho, lb, ub, xn, yn, dx, dy, diff, work)
