"""Typed Python front of the C ABI: torch CUDA tensors in, kernels launched on torch's current stream.

One function per entry point of include/gradflow_b200.h.  Tensors must be contiguous, float64 (int32 /
uint8 where stated) and live on the current CUDA device.  Nothing here computes on the CPU.
"""

from __future__ import annotations

from typing import Optional

import torch

from . import native
from .native import ptr

LAUNCHES = 0  # number of kernel-launching ABI calls (bench.py reports it as gpu_launches evidence)


class WorkList:
    """Optional instance list for the batch dimension (see the header's conventions)."""

    __slots__ = ("list", "count_dev", "nwork")

    def __init__(self, list_t: Optional[torch.Tensor], count_dev: Optional[torch.Tensor], nwork: int):
        self.list = list_t
        self.count_dev = count_dev
        self.nwork = int(nwork)

    @staticmethod
    def all(B: int) -> "WorkList":
        return WorkList(None, None, B)


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _w(work: WorkList):
    return (ptr(work.list), ptr(work.count_dev), work.nwork, _stream())


def _call(name, *args, launches: int = 1):
    global LAUNCHES
    LAUNCHES += launches
    native.call(name, *args)


def qp_eval(H, A, g, b, x, grad, cons, obj, work: WorkList):
    B, n = x.shape
    m = 0 if A is None else A.shape[1]
    _call("gf_qp_eval", B, n, m, ptr(H), ptr(A), ptr(g), ptr(b), ptr(x), ptr(grad), ptr(cons), ptr(obj), *_w(work))


def rosen_eval(a, b, x, grad, obj, work: WorkList):
    B, n = x.shape
    _call("gf_rosen_eval", B, n, ptr(a), ptr(b), ptr(x), ptr(grad), ptr(obj), *_w(work))


def rosen_hess(a, b, x, H, work: WorkList):
    B, n = x.shape
    _call("gf_rosen_hess", B, n, ptr(a), ptr(b), ptr(x), ptr(H), *_w(work))


def ocp_eval(S, nx, nu, h, A, Bm, Q, R, xinit, z, grad, cons, obj, work: WorkList):
    _call("gf_ocp_eval", z.shape[0], S, nx, nu, h, ptr(A), ptr(Bm), ptr(Q), ptr(R), ptr(xinit), ptr(z), ptr(grad),
          ptr(cons), ptr(obj), *_w(work))


def ocp_jac(S, nx, nu, h, A, Bm, z, J, work: WorkList):
    _call("gf_ocp_jac", z.shape[0], S, nx, nu, h, ptr(A), ptr(Bm), ptr(z), ptr(J), *_w(work))


def ocp_hess(S, nx, nu, c1, Q, R, z, y, H, work: WorkList):
    _call("gf_ocp_hess", z.shape[0], S, nx, nu, c1, ptr(Q), ptr(R), ptr(z), ptr(y), ptr(H), *_w(work))


def aug_lag_grad(J, grad, cons, y, rho, dL, jty, jtc, work: WorkList):
    B, n = grad.shape
    m = 0 if cons is None else cons.shape[1]
    _call("gf_aug_lag_grad", B, n, m, ptr(J), ptr(grad), ptr(cons), ptr(y), ptr(rho), ptr(dL), ptr(jty), ptr(jtc),
          *_w(work))


def residual(x, y, x0, y0, dL, cons, lb, ub, dt, scaled: bool, active_mode: int, active, F, nrm, work: WorkList,
             tau=None):
    B, n = x.shape
    m = 0 if y is None else y.shape[1]
    if tau is not None:
        _call("gf_residual_tau", B, n, m, ptr(x), ptr(y), ptr(x0), ptr(y0), ptr(dL), ptr(cons), ptr(lb), ptr(ub),
              ptr(dt), ptr(tau), 1 if scaled else 0, active_mode, ptr(active), ptr(F), ptr(nrm), *_w(work))
        return
    _call("gf_residual", B, n, m, ptr(x), ptr(y), ptr(x0), ptr(y0), ptr(dL), ptr(cons), ptr(lb), ptr(ub), ptr(dt),
          1 if scaled else 0, active_mode, ptr(active), ptr(F), ptr(nrm), *_w(work))


def index_sets(active, m: int, perm, nI, Nvec, work: WorkList):
    B, n = active.shape
    _call("gf_index_sets", B, n, m, ptr(active), ptr(perm), ptr(nI), ptr(Nvec), *_w(work))


def kkt_assemble(H, J, perm, nI, dt, rho, K, pad: int, lower_only: bool, work: WorkList):
    B, n, _ = H.shape
    m = 0 if J is None else J.shape[1]
    ld = K.shape[1]
    _call("gf_kkt_assemble", B, n, m, ld, pad, 1 if lower_only else 0, ptr(H), ptr(J), ptr(perm), ptr(nI), ptr(dt),
          ptr(rho), ptr(K), *_w(work))


def kkt_rhs(H, J, perm, nI, F, dt, rho, rhs, work: WorkList):
    B, n, _ = H.shape
    m = 0 if J is None else J.shape[1]
    ld = rhs.shape[1]
    _call("gf_kkt_rhs", B, n, m, ld, ptr(H), ptr(J), ptr(perm), ptr(nI), ptr(F), ptr(dt), ptr(rho), ptr(rhs),
          *_w(work))


FORM_SYMMETRIC, FORM_ASYMMETRIC, FORM_EXTENDED, FORM_STANDARD, FORM_SCALED_DERIV = 0, 1, 2, 3, 4  # GF_FORM_* of the header


def kkt_assemble_full(H, J, perm, nI, active, dt, rho, K, form: int, work: WorkList):
    B, n, _ = H.shape
    m = 0 if J is None else J.shape[1]
    _call("gf_kkt_assemble_full", B, n, m, K.shape[1], form, ptr(H), ptr(J), ptr(perm), ptr(nI), ptr(active), ptr(dt),
          ptr(rho), ptr(K), *_w(work))


def kkt_rhs_full(n: int, m: int, perm, nI, active, F, dt, rho, rhs, form: int, work: WorkList):
    _call("gf_kkt_rhs_full", rhs.shape[0], n, m, rhs.shape[1], form, ptr(perm), ptr(nI), ptr(active), ptr(F), ptr(dt),
          ptr(rho), ptr(rhs), *_w(work))


def lu_factor(K, Nmax: int, Nvec, piv, info, work: WorkList):
    B, ld, _ = K.shape
    _call("gf_lu_factor", B, ld, Nmax, ptr(Nvec), ptr(K), ptr(piv), ptr(info), *_w(work))


def lu_solve(K, Nmax: int, Nvec, piv, rhs, trans: bool, work: WorkList):
    B, ld, _ = K.shape
    _call("gf_lu_solve", B, ld, Nmax, ptr(Nvec), ptr(K), ptr(piv), ptr(rhs), rhs.shape[1], 1 if trans else 0,
          *_w(work))


def pareto_update(grad, cons, jty, jtc, phase, status, opt_tol, local_infeas_tol, rho, work: WorkList):
    """ParetoDecrease.update (penalty.py:136-168) on the committed iterates of the instances that just accepted."""
    B, n = grad.shape
    m = 0 if cons is None else cons.shape[1]
    _call("gf_pareto_update", B, n, m, ptr(grad), ptr(cons), ptr(jty), ptr(jtc), ptr(phase), ptr(status), opt_tol,
          local_infeas_tol, ptr(rho), *_w(work))


def filter_update(kind: int, n: int, m: int, phase, status, om, cm, dm, of, cf, df, rho, rho_pen, filt, nfilt, overflow,
                  work: WorkList):
    """ObjectivePenaltyFilter (kind 0) / LagrangianPenaltyFilter (kind 1).update + the veto of solver.py:357-378."""
    B = phase.shape[0]
    _call("gf_filter_update", B, n, m, kind, ptr(phase), ptr(status), ptr(om), ptr(cm), ptr(dm), ptr(of), ptr(cf), ptr(df),
          ptr(rho), ptr(rho_pen), ptr(filt), ptr(nfilt), filt.shape[1], ptr(overflow), *_w(work))


def rosen_fused_solve(a, b, lb, ub, x, grad, obj, lamb, err_sum, status, iters, accepted, newton_steps, total_res, active,
                      params12, iteration_limit, max_outer: int, fresh: bool, work: WorkList):
    """Whole Solver.solve of small chained-Rosenbrock instances, one warp per instance (see the header)."""
    import ctypes

    B, n = x.shape
    arr = (ctypes.c_double * 12)(*[float(v) for v in params12])
    _call("gf_rosen_fused_solve", B, n, ptr(a), ptr(b), ptr(lb), ptr(ub), ptr(x), ptr(grad), ptr(obj), ptr(lamb),
          ptr(err_sum), ptr(status), ptr(iters), ptr(accepted), ptr(newton_steps), ptr(total_res), ptr(active),
          ctypes.cast(arr, ctypes.c_void_p), -1 if iteration_limit is None else int(iteration_limit), int(max_outer),
          1 if fresh else 0, *_w(work))


def hess_rho(H, J, rho, out, work: WorkList):
    """out = H + rho J'J (aug_lag_deriv_xx(rho) of the Standard formulation, iterate.py:103-110)."""
    B, m, n = J.shape
    _call("gf_hess_rho", B, n, m, ptr(H), ptr(J), ptr(rho), ptr(out), *_w(work))


def krylov_scratch_rows(minres: bool, restart: int = 20) -> int:
    return native.load().gf_krylov_scratch_rows(1 if minres else 0, restart)


def gmres_solve(K, Nmax: int, Nvec, rhs, x0, x0_mask, trans: bool, scratch, info, iters, work: WorkList,
                restart: int = 20, rtol: float = 1e-5, atol: float = 1e-8):
    """GMRESSolver.solve (gmres_solver.py:12-35): rhs <- solution; x0 / x0_mask = the start vector (see the header)."""
    B, ld, _ = K.shape
    nmask = 0 if x0_mask is None else x0_mask.shape[1]
    _call("gf_gmres_solve", B, ld, Nmax, ptr(Nvec), ptr(K), ptr(rhs), rhs.shape[1], ptr(x0), ptr(x0_mask), nmask,
          1 if trans else 0, restart, rtol, atol, ptr(scratch), ptr(info), ptr(iters), *_w(work))


def minres_solve(K, Nmax: int, Nvec, rhs, x0, scratch, info, iters, work: WorkList, rtol: float = 1e-5):
    """MINRESSolver.solve (minres_solver.py:12-24): rhs <- solution."""
    B, ld, _ = K.shape
    _call("gf_minres_solve", B, ld, Nmax, ptr(Nvec), ptr(K), ptr(rhs), rhs.shape[1], ptr(x0), rtol, ptr(scratch),
          ptr(info), ptr(iters), *_w(work))


def ldlt_factor(K, Nmax: int, Nvec, dvec, info, nneg, npos_expected, work: WorkList):
    B, ld, _ = K.shape
    nblk = (Nmax + 63) // 64
    _call("gf_ldlt_factor", B, ld, Nmax, ptr(Nvec), ptr(K), ptr(dvec), ptr(info), ptr(nneg), ptr(npos_expected),
          *_w(work), launches=nblk * (2 if work.nwork >= 1024 and nblk > 2 else 1))


def kkt_ldlt_factor(H, J, perm, nI, dt, rho, Nvec, K, dvec, info, nneg, work: WorkList):
    """gf_kkt_assemble (lower, padded) fused into gf_ldlt_factor."""
    B, n, _ = H.shape
    m = 0 if J is None else J.shape[1]
    ld = K.shape[1]
    nblk = (n + m + 63) // 64
    _call("gf_kkt_ldlt_factor", B, n, m, ld, ptr(H), ptr(J), ptr(perm), ptr(nI), ptr(dt), ptr(rho), ptr(Nvec), ptr(K),
          ptr(dvec), ptr(info), ptr(nneg), *_w(work), launches=nblk * (2 if work.nwork >= 1024 and nblk > 2 else 1))


def ldlt_solve(K, Nmax: int, Nvec, rhs, work: WorkList):
    B, ld, _ = K.shape
    _call("gf_ldlt_solve", B, ld, Nmax, ptr(Nvec), ptr(K), ptr(rhs), rhs.shape[1], *_w(work))


def step_finish(xbase, ybase, sol, perm, nI, F, dt, rho, lb, ub, xn, yn, dx, dy, diff, work: WorkList):
    B, n = xbase.shape
    m = 0 if ybase is None else ybase.shape[1]
    ld = sol.shape[1]
    _call("gf_step_finish", B, n, m, ld, ptr(xbase), ptr(ybase), ptr(sol), ptr(perm), ptr(nI), ptr(F), ptr(dt),
          ptr(rho), ptr(lb), ptr(ub), ptr(xn), ptr(yn), ptr(dx), ptr(dy), ptr(diff), *_w(work))


def check_terminate(x, grad, cons, jty, jtc, obj, lb, ub, opt_tol, active_tol, local_infeas_tol, obj_lower_limit,
                    iteration_limit, iters, status, total_res, work: WorkList):
    B, n = x.shape
    m = 0 if cons is None else cons.shape[1]
    _call("gf_check_terminate", B, n, m, ptr(x), ptr(grad), ptr(cons), ptr(jty), ptr(jtc), ptr(obj), ptr(lb),
          ptr(ub), opt_tol, active_tol, local_infeas_tol, obj_lower_limit,
          -1 if iteration_limit is None else int(iteration_limit), ptr(iters), ptr(status), ptr(total_res), *_w(work))


def dr_first(status, info, dt, mid_norm, diff1, newton_tol, lamb_red, lamb_min, phase, lamb_next):
    _call("gf_dr_first", status.shape[0], ptr(status), ptr(info), ptr(dt), ptr(mid_norm), ptr(diff1), newton_tol,
          lamb_red, lamb_min, ptr(phase), ptr(lamb_next), _stream())


def dr_second(dt, diff1, diff2, theta_max, log_theta_ref, K_P, K_I, lamb_min, lamb_inc, err_sum, phase, lamb_next,
              theta):
    _call("gf_dr_second", dt.shape[0], ptr(dt), ptr(diff1), ptr(diff2), theta_max, log_theta_ref, K_P, K_I, lamb_min,
          lamb_inc, ptr(err_sum), ptr(phase), ptr(lamb_next), ptr(theta), _stream())


def count_newton_steps(phase, nsteps):
    _call("gf_count_newton_steps", phase.shape[0], ptr(phase), ptr(nsteps), _stream())


def single_control(fixed, status, info, dt, mid_norm, orig_norm, prm, err_sum, phase, lamb_next, theta, nsteps):
    _call("gf_single_control", status.shape[0], 1 if fixed else 0, ptr(status), ptr(info), ptr(dt), ptr(mid_norm),
          ptr(orig_norm), prm.newton_tol, prm.theta_max, prm.log_theta_ref, prm.K_P, prm.K_I, prm.lamb_red, prm.lamb_min,
          prm.lamb_inc, prm.lamb_init, ptr(err_sum), ptr(phase), ptr(lamb_next), ptr(theta), ptr(nsteps), _stream())


def exact_control(mode, it, last, status, info, dt, val, orig_norm, newton_tol, rate_bound, curr, live, phase, lamb_next,
                  nsteps):
    _call("gf_exact_control", status.shape[0], mode, it, 1 if last else 0, ptr(status), ptr(info), ptr(dt), ptr(val),
          ptr(orig_norm), newton_tol, rate_bound, ptr(curr), ptr(live), ptr(phase), ptr(lamb_next), ptr(nsteps), _stream())


def commit(phase, lamb_next, lamb_max, dual_norm_update, mid, fin, cur, lamb, rho, iters, accepted, status):
    """mid / fin / cur: tuples (x, y, grad, cons, obj)."""
    B, n = cur[0].shape
    m = 0 if cur[1] is None else cur[1].shape[1]
    _call("gf_commit", B, n, m, ptr(phase), ptr(lamb_next), lamb_max, int(dual_norm_update),
          *[ptr(t) for t in mid], *[ptr(t) for t in fin], *[ptr(t) for t in cur], ptr(lamb), ptr(rho), ptr(iters),
          ptr(accepted), ptr(status), _stream())


def build_worklist(key, lo: int, hi: int, out: WorkList, parent: Optional[WorkList] = None, invert: bool = False):
    """out.list / out.count_dev <- ordered { b in parent : (lo <= key[b] <= hi) != invert }."""
    plist = None if parent is None else parent.list
    pcount = None if parent is None else parent.count_dev
    _call("gf_build_worklist", key.shape[0], ptr(key), lo, hi, 1 if invert else 0, ptr(plist), ptr(pcount),
          ptr(out.list), ptr(out.count_dev), _stream())


def dt_from_lamb(lamb, dt):
    _call("gf_dt_from_lamb", lamb.shape[0], ptr(lamb), ptr(dt), _stream())


def merit_grad(H, J, F, active, dt, rho, dx, dy, res, inner, work: WorkList):
    B, n, _ = H.shape
    m = 0 if J is None else J.shape[1]
    _call("gf_merit_grad", B, n, m, ptr(H), ptr(J), ptr(F), ptr(active), ptr(dt), ptr(rho), ptr(dx), ptr(dy),
          ptr(res), ptr(inner), *_w(work))


def ls_begin(res, newton_tol, state, alpha, trials, work: WorkList):
    _call("gf_ls_begin", res.shape[0], ptr(res), newton_tol, ptr(state), ptr(alpha), ptr(trials), *_w(work))


def ls_trial(x, y, dx, dy, alpha, xt, yt, work: WorkList):
    B, n = x.shape
    m = 0 if y is None else y.shape[1]
    _call("gf_ls_trial", B, n, m, ptr(x), ptr(y), ptr(dx), ptr(dy), ptr(alpha), ptr(xt), ptr(yt), *_w(work))


def armijo_residual(xt, yt, x0, y0, dL, cons, lb, ub, dt, res, inner, newton_tol, max_trials, alpha, trials, state,
                    next_res, work: WorkList):
    B, n = xt.shape
    m = 0 if yt is None else yt.shape[1]
    _call("gf_armijo_residual", B, n, m, ptr(xt), ptr(yt), ptr(x0), ptr(y0), ptr(dL), ptr(cons), ptr(lb), ptr(ub),
          ptr(dt), ptr(res), ptr(inner), newton_tol, max_trials, ptr(alpha), ptr(trials), ptr(state), ptr(next_res),
          *_w(work))


SYM_BLOCK = 64  # row-block height of the symmetric transfer


def ldexp(src, out, work: WorkList, rw=None, sr: int = 0, cw=None, sc: int = 0, ow=None, so: int = 0):
    """out[b][r][c] = ldexp(src[b][r][c], sr rw[b][r] + sc cw[b][c] + so ow[b]); src [B, cols] or [B, rows, cols]."""
    B = src.shape[0]
    rows, cols = (1, src.shape[1]) if src.dim() == 2 else (src.shape[1], src.shape[2])
    if src.dim() == 1:
        rows, cols = 1, 1
    _call("gf_ldexp", B, rows, cols, ptr(src), ptr(rw), sr, ptr(cw), sc, ptr(ow), so, ptr(out), *_w(work))


def h2d_sym_lower(dst, src_host, cnt: int):
    """dst[:cnt] (device, [*, n, n]) <- lower block triangle of the pinned host tensor src_host[:cnt]."""
    n = dst.shape[1]
    assert src_host.is_pinned() and src_host.is_contiguous() and dst.is_contiguous()
    _call("gf_h2d_sym_lower", ptr(dst), src_host.data_ptr(), cnt, n, SYM_BLOCK, _stream(), launches=0)


def h2d_sym_lower_bytes(cnt: int, n: int) -> int:
    tot = 0
    for r0 in range(0, n, SYM_BLOCK):
        r1 = min(n, r0 + SYM_BLOCK)
        tot += r1 * (r1 - r0) * 8
    return tot * cnt


def symmetrize_lower(H, cnt: int):
    """Rebuild the blocks above the diagonal of H[:cnt] from the transferred lower block triangle."""
    _call("gf_symmetrize_lower", ptr(H), cnt, H.shape[1], SYM_BLOCK, _stream())


def band_assemble(H, J, active, order, bw: int, dt, rho, Kband, work: WorkList):
    B, n, _ = H.shape
    m = 0 if J is None else J.shape[1]
    _call("gf_band_assemble", B, n, m, bw, ptr(H), ptr(J), ptr(active), ptr(order), ptr(dt), ptr(rho), ptr(Kband),
          *_w(work))


def band_factor(Kband, bw: int, info, nneg, work: WorkList):
    B, N, _ = Kband.shape
    _call("gf_band_factor", B, N, bw, ptr(Kband), ptr(info), ptr(nneg), *_w(work))


def band_solve(Kband, bw: int, v, work: WorkList):
    B, N, _ = Kband.shape
    _call("gf_band_solve", B, N, bw, ptr(Kband), ptr(v), *_w(work))


def band_permute(perm, nI, pos, stdv, bandv, to_band: bool, m: int, work: WorkList):
    B, n = perm.shape
    _call("gf_band_permute", B, n, m, stdv.shape[1], ptr(perm), ptr(nI), ptr(pos), ptr(stdv), ptr(bandv),
          1 if to_band else 0, *_w(work))


# ---- stage-structured KKT systems (cfg4): compact Jacobian Jc [B, S*nx, nx + w], diagonal Hessian Hd [B, n] ----------
def ocp_jac_banded(S, nx, nu, h, A, Bm, z, Jc, diag_only: bool, work: WorkList):
    _call("gf_ocp_jac_banded", z.shape[0], S, nx, nu, h, ptr(A), ptr(Bm), ptr(z), ptr(Jc), 1 if diag_only else 0,
          *_w(work))


def ocp_hess_diag(S, nx, nu, c1, Q, R, z, y, Hd, work: WorkList):
    _call("gf_ocp_hess_diag", z.shape[0], S, nx, nu, c1, ptr(Q), ptr(R), ptr(z), ptr(y), ptr(Hd), *_w(work))


def stage_aug_lag_grad(S, nx, nu, Jc, grad, cons, y, rho, dL, jty, jtc, work: WorkList):
    _call("gf_stage_aug_lag_grad", grad.shape[0], S, nx, nu, ptr(Jc), ptr(grad), ptr(cons), ptr(y), ptr(rho), ptr(dL),
          ptr(jty), ptr(jtc), *_w(work))


def stage_kkt_factor(S, nx, nu, Jc, Hd, active, dt, rho, Tinv, Pf, Qf, info, nneg, work: WorkList):
    _call("gf_stage_kkt_factor", Hd.shape[0], S, nx, nu, ptr(Jc), ptr(Hd), ptr(active), ptr(dt), ptr(rho), ptr(Tinv),
          ptr(Pf), ptr(Qf), ptr(info), ptr(nneg), *_w(work))


def stage_kkt_solve(S, nx, nu, Jc, Hd, active, F, dt, rho, Tinv, Pf, Qf, sol, work: WorkList):
    _call("gf_stage_kkt_solve", Hd.shape[0], S, nx, nu, ptr(Jc), ptr(Hd), ptr(active), ptr(F), ptr(dt), ptr(rho),
          ptr(Tinv), ptr(Pf), ptr(Qf), ptr(sol), sol.shape[1], *_w(work))
