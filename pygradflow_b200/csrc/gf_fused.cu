// Fused persistent solver for small bound-constrained instances (cfg2: chained Rosenbrock, n <= 64, m = 0).
//
// One WARP runs Solver.solve for one instance from start to finish: termination test, residual + active set, Hessian,
// reduced KKT system, factorisation, two substitutions, step finish, the DistanceRatio controller with its log-PI
// update and the commit -- the per-instance state machine of pygradflow_b200/solver.py (which restates
// solver.py:180-205,305-380, distance_ratio_control.py:18-78, controller.py:29-77, step_control.py:64-107,
// newton.py:35-60, symmetric_step_solver.py:27-164, scaled_step_solver.py:38-107, step_solver.py:16-63) with every
// value in registers / shared memory and no lock-step: an instance that needs 30 000 outer iterations no longer makes
// 4095 others wait at 40 kernel launches per iteration.  All 4096 warps of cfg2 are resident at once (28 per SM).
//
// The arithmetic of every stage is the one of the stand-alone kernels (gf_eval.cu, gf_step.cu), operation by operation
// and with the same reduction trees, except the linear solve: the Hessian of this family is tridiagonal, hence so is
// the reduced matrix H_II + lamb I, and it is factorised as such -- LAPACK dgttrf / dgttrs (row interchanges between
// neighbours, second super-diagonal) per maximal run of consecutive inactive variables, the runs eliminated in
// parallel by the lanes that own their first element.  The reference factorises the same sparse matrix with SuperLU
// (lu_solver.py:14); both are backward-stable partial-pivoting eliminations, results agree to rounding.
#include "gf_common.cuh"
#include "../../include/gradflow_b200.h"

namespace {

constexpr int FW = 4;          // warps (instances) per CTA
constexpr int FN = 64;         // maximum number of variables
constexpr int NPL = FN / 32;   // elements per lane: lane l owns l and l + 32

struct FusedPrm {
    double opt_tol, active_tol, obj_lower_limit, newton_tol, lamb_red, lamb_min, lamb_max, lamb_inc, theta_max,
        log_theta_ref, K_P, K_I;
    int iteration_limit;
};

struct WarpSmem {
    double xs[FN];    // point whose neighbours are read (x, mid, fin in turn)
    double a[FN];     // family coefficients a_i, b_i (i < n - 1)
    double b[FN];
    double ho[FN];    // H_{i,i+1}
    double F[FN];     // dt * F_i of the active variables (b0), for the neighbours' right-hand sides
    double D[FN], DL[FN], DU[FN], DU2[FN], RD[FN], R[FN];
    unsigned char P[FN];
};

// block_sum of a 64-thread CTA (gf_common.cuh): xor tree inside each warp, then the two partials added.
__device__ __forceinline__ double tree_sum(const double (&v)[NPL]) {
    double r = warp_sum(v[0]);
#pragma unroll
    for (int k = 1; k < NPL; ++k) r = r + warp_sum(v[k]);
    return r;
}
__device__ __forceinline__ double tree_max(const double (&v)[NPL]) {
    double r = warp_max(v[0]);
#pragma unroll
    for (int k = 1; k < NPL; ++k) r = fmax(r, warp_max(v[k]));
    return r;
}

// rosen_eval_kernel (gf_eval.cu) on the point in S.xs: gradient per owned element, objective.
__device__ __forceinline__ void rosen_eval(const WarpSmem& S, int n, int lane, double (&g)[NPL], double& obj) {
    double op[NPL];
#pragma unroll
    for (int k = 0; k < NPL; ++k) {
        const int i = lane + 32 * k;
        double gi = 0.0, o = 0.0;
        if (i < n) {
            const double xi = S.xs[i];
            if (i < n - 1) {
                const double r = __dsub_rn(S.xs[i + 1], __dmul_rn(xi, xi));
                const double d = __dsub_rn(S.a[i], xi);
                const double t1 = __dmul_rn(__dmul_rn(__dmul_rn(-4.0, S.b[i]), r), xi);
                gi = __dadd_rn(gi, __dsub_rn(t1, __dmul_rn(2.0, d)));
                o = __dadd_rn(__dmul_rn(__dmul_rn(S.b[i], r), r), __dmul_rn(d, d));
            }
            if (i > 0) {
                const double xm = S.xs[i - 1];
                const double rm = __dsub_rn(xi, __dmul_rn(xm, xm));
                gi = __dadd_rn(gi, __dmul_rn(__dmul_rn(2.0, S.b[i - 1]), rm));
            }
        }
        g[k] = gi;
        op[k] = o;
    }
    obj = tree_sum(op);
}

// rosen_hess_kernel on S.xs: main diagonal per owned element, H_{i,i+1} into S.ho.
__device__ __forceinline__ void rosen_hess(WarpSmem& S, int n, int lane, double (&hd)[NPL]) {
#pragma unroll
    for (int k = 0; k < NPL; ++k) {
        const int i = lane + 32 * k;
        double main = 0.0;
        if (i < n) {
            const double xi = S.xs[i];
            double off = 0.0;
            if (i < n - 1) {
                const double r = __dsub_rn(S.xs[i + 1], __dmul_rn(xi, xi));
                const double t = __dsub_rn(__dmul_rn(__dmul_rn(8.0, S.b[i]), __dmul_rn(xi, xi)),
                                           __dmul_rn(__dmul_rn(4.0, S.b[i]), r));
                main = __dadd_rn(main, __dadd_rn(t, 2.0));
                off = __dmul_rn(__dmul_rn(-4.0, S.b[i]), xi);
            }
            if (i > 0) main = __dadd_rn(main, __dmul_rn(2.0, S.b[i - 1]));
            S.ho[i] = off;
        }
        hd[k] = main;
    }
}

// dgttrf on the run [s, s + L): D / DL / DU hold the diagonals (DL[i] = DU[i] = coupling of i and i + 1 on entry).
// Returns false on an exactly zero or non-finite pivot (LUSolver raises LinearSolverError, lu_solver.py:15-17).
__device__ __forceinline__ bool tri_factor_run(WarpSmem& S, int s, int L) {
    double dcur = S.D[s];
    double ucur = L > 1 ? S.DU[s] : 0.0;
    bool ok = true;
    for (int i = s; i < s + L - 1; ++i) {
        const double l = S.DL[i];
        double dnext = S.D[i + 1];
        double unext = (i + 1 < s + L - 1) ? S.DU[i + 1] : 0.0;
        if (fabs(dcur) >= fabs(l)) {
            double f = l;
            if (dcur != 0.0) {
                f = l / dcur;
                dnext = dnext - f * ucur;
            }
            S.DL[i] = f;
            S.D[i] = dcur;
            S.DU[i] = ucur;
            S.DU2[i] = 0.0;
            S.P[i] = 0;
            ok = ok && dcur != 0.0 && isfinite(dcur);
            S.RD[i] = 1.0 / dcur;
        } else {  // interchange rows i and i + 1
            const double f = dcur / l;
            S.D[i] = l;
            S.RD[i] = 1.0 / l;
            S.DL[i] = f;
            S.DU[i] = dnext;
            dnext = ucur - f * dnext;
            S.DU2[i] = unext;
            unext = -f * unext;
            S.P[i] = 1;
            ok = ok && isfinite(l);
        }
        dcur = dnext;
        ucur = unext;
    }
    S.D[s + L - 1] = dcur;
    S.RD[s + L - 1] = 1.0 / dcur;
    return ok && dcur != 0.0 && isfinite(dcur);
}

// dgttrs (no transpose) on the run: S.R[s .. s + L) <- solution.
__device__ __forceinline__ void tri_solve_run(WarpSmem& S, int s, int L) {
    double bcur = S.R[s];
    for (int i = s; i < s + L - 1; ++i) {
        double bn = S.R[i + 1];
        if (S.P[i] == 0) {
            S.R[i] = bcur;
            bn = bn - S.DL[i] * bcur;
        } else {
            S.R[i] = bn;
            bn = bcur - S.DL[i] * bn;
        }
        bcur = bn;
    }
    const int e = s + L - 1;
    double x1 = bcur * S.RD[e];
    S.R[e] = x1;
    if (L > 1) {
        double x0 = (S.R[e - 1] - S.DU[e - 1] * x1) * S.RD[e - 1];
        S.R[e - 1] = x0;
        for (int i = e - 2; i >= s; --i) {
            const double v = (S.R[i] - S.DU[i] * x0 - S.DU2[i] * x1) * S.RD[i];
            S.R[i] = v;
            x1 = x0;
            x0 = v;
        }
    }
}

__global__ void __launch_bounds__(FW * 32) fused_rosen_kernel(
    int n, const double* __restrict__ a_all, const double* __restrict__ b_all, const double* __restrict__ lb_all,
    const double* __restrict__ ub_all, double* __restrict__ x_all, double* __restrict__ grad_all,
    double* __restrict__ obj_all, double* __restrict__ lamb_all, double* __restrict__ errsum_all,
    int32_t* __restrict__ status_all, int32_t* __restrict__ iters_all, int32_t* __restrict__ accepted_all,
    int32_t* __restrict__ nsteps_all, double* __restrict__ totres_all, uint8_t* __restrict__ active_all, FusedPrm prm,
    int max_outer, int fresh, GfWork work, int nwork) {
    __shared__ WarpSmem smem[FW];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slot = blockIdx.x * FW + wid;
    if (slot >= nwork) return;
    const int b = gf_instance(work, slot);
    if (b < 0) return;
    if (status_all[b] != 0) return;
    WarpSmem& S = smem[wid];
    const unsigned FULL = 0xffffffffu;

    double x[NPL], g[NPL], lb[NPL], ub[NPL];
#pragma unroll
    for (int k = 0; k < NPL; ++k) {
        const int i = lane + 32 * k;
        const bool in = i < n;
        x[k] = in ? x_all[(size_t)b * n + i] : 0.0;
        lb[k] = in ? lb_all[(size_t)b * n + i] : 0.0;
        ub[k] = in ? ub_all[(size_t)b * n + i] : 0.0;
        S.a[i] = (i < n - 1) ? a_all[(size_t)b * (n - 1) + i] : 0.0;
        S.b[i] = (i < n - 1) ? b_all[(size_t)b * (n - 1) + i] : 0.0;
        S.xs[i] = x[k];
    }
    __syncwarp();
    double obj;
    if (fresh) {
        rosen_eval(S, n, lane, g, obj);
    } else {
#pragma unroll
        for (int k = 0; k < NPL; ++k) g[k] = (lane + 32 * k < n) ? grad_all[(size_t)b * n + lane + 32 * k] : 0.0;
        obj = obj_all[b];
    }
    double lamb_cur = lamb_all[b];
    double err_sum = errsum_all[b];
    int iters = iters_all[b], accepted = accepted_all[b], nsteps = nsteps_all[b];
    int status = 0;
    double total_res = 0.0;
    unsigned long long amask = 0ull;  // active set of the last factorisation (bit i)

    for (int outer = 0; outer < max_outer; ++outer) {
        // ---- Solver._check_terminate (terminate_kernel with m = 0) -------------------------------------------
        {
            double bv[NPL], st[NPL];
#pragma unroll
            for (int k = 0; k < NPL; ++k) {
                bv[k] = 0.0;
                st[k] = 0.0;
                if (lane + 32 * k < n) {
                    const double xi = x[k], l = lb[k], u = ub[k];
                    bv[k] = fmax(fmax(l - xi, 0.0), fmax(xi - u, 0.0));
                    const bool atl = fabs(xi - l) <= prm.active_tol;
                    const bool atu = fabs(u - xi) <= prm.active_tol;
                    const bool both = atl && atu, lo = atl && !both, up = atu && !both;
                    const double gj = __dadd_rn(g[k], 0.0);
                    const double r = -gj;
                    double d = 0.0;
                    if (up) d = fmax(r, 0.0);
                    if (lo) d = fmin(r, 0.0);
                    if (both) d = r;
                    st[k] = fabs(__dadd_rn(gj, d));
                }
            }
            const double bvm = tree_max(bv), stm = tree_max(st);
            total_res = fmax(0.0, fmax(bvm, stm));
            if (prm.iteration_limit >= 0 && iters >= prm.iteration_limit) status = GF_STATUS_ITERATION_LIMIT;
            else if (total_res <= prm.opt_tol) status = GF_STATUS_OPTIMAL;
            else if (obj <= prm.obj_lower_limit && bvm <= prm.opt_tol) status = GF_STATUS_UNBOUNDED;
            if (status != 0) break;
        }
        const double dtb = 1.0 / lamb_cur;   // gf_dt_from_lamb
        const double lamb = 1.0 / dtb;       // implicit_func.py:212
        // ---- first Newton step: residual + active set at the current iterate (residual_kernel, scaled, mode 0) ----
        double F[NPL];
        bool act[NPL];
#pragma unroll
        for (int k = 0; k < NPL; ++k) {
            const int i = lane + 32 * k;
            const double p = __dsub_rn(__dmul_rn(lamb, x[k]), g[k]);
            const double lo = __dmul_rn(lamb, lb[k]), hi = __dmul_rn(lamb, ub[k]);
            const double xsc = __dmul_rn(lamb, x[k]);
            act[k] = (i < n) && ((p < lo - GF_ACTIVE_SLACK) || (p > hi + GF_ACTIVE_SLACK));
            const double proj = act[k] ? fmin(fmax(p, lo), hi) : p;
            F[k] = __dsub_rn(xsc, proj);
        }
        amask = 0ull;
#pragma unroll
        for (int k = 0; k < NPL; ++k) amask |= (unsigned long long)__ballot_sync(FULL, act[k]) << (32 * k);
        // variables beyond n count as active (identity rows): they end every run
        const unsigned long long valid = n >= 64 ? ~0ull : ((1ull << n) - 1ull);
        const unsigned long long inact = ~amask & valid;
        // ---- Hessian at the current iterate, reduced system K = H_II + lamb I (kkt_assemble_kernel) ----
        double hd[NPL];
        rosen_hess(S, n, lane, hd);  // S.xs holds the current iterate here
#pragma unroll
        for (int k = 0; k < NPL; ++k) {
            const int i = lane + 32 * k;
            S.D[i] = act[k] || i >= n ? 1.0 : __dadd_rn(hd[k], lamb);
        }
        __syncwarp();
#pragma unroll
        for (int k = 0; k < NPL; ++k) {
            const int i = lane + 32 * k;
            const bool coupled = i + 1 < n && ((inact >> i) & 3ull) == 3ull;
            const double e = coupled ? S.ho[i] : 0.0;
            S.DL[i] = e;
            S.DU[i] = e;
        }
        __syncwarp();
        // ---- factorisation: one lane per run of consecutive inactive variables ----
        bool ok = true;
#pragma unroll
        for (int k = 0; k < NPL; ++k) {
            const int i = lane + 32 * k;
            const bool start = ((inact >> i) & 1ull) && (i == 0 || !((inact >> (i - 1)) & 1ull));
            if (start) {
                const unsigned long long rest = ~inact >> i;  // first non-inactive position at or after i
                const int L = rest ? __ffsll((long long)rest) - 1 : 64 - i;
                ok = tri_factor_run(S, i, L) && ok;
            }
        }
        const bool failed = __any_sync(FULL, !ok);
        __syncwarp();

        int phase;
        double lamb_next = lamb;
        double xm[NPL], gm[NPL], xf[NPL], gf[NPL];
        double objm = 0.0, objf = 0.0;
        double diff1 = 0.0, diff2 = 0.0;
#pragma unroll
        for (int k = 0; k < NPL; ++k) { xm[k] = x[k]; gm[k] = g[k]; xf[k] = x[k]; gf[k] = g[k]; }

        // One Newton step of the frozen system from `base` with residual Fv: rhs, substitution, step finish.
        auto newton_step = [&](const double (&base)[NPL], const double (&Fv)[NPL], double (&out)[NPL]) -> double {
            // b0 = dt F[A] for the neighbours (kkt_rhs_kernel)
#pragma unroll
            for (int k = 0; k < NPL; ++k) S.F[lane + 32 * k] = __dmul_rn(dtb, Fv[k]);
            __syncwarp();
#pragma unroll
            for (int k = 0; k < NPL; ++k) {
                const int i = lane + 32 * k;
                double r = 0.0;
                if (i < n && !act[k]) {
                    double acc = 0.0;
                    if (i > 0 && ((amask >> (i - 1)) & 1ull)) acc = __dadd_rn(acc, __dmul_rn(S.ho[i - 1], S.F[i - 1]));
                    if (i + 1 < n && ((amask >> (i + 1)) & 1ull)) acc = __dadd_rn(acc, __dmul_rn(S.ho[i], S.F[i + 1]));
                    r = __dsub_rn(Fv[k], acc);
                }
                S.R[i] = r;
            }
            __syncwarp();
#pragma unroll
            for (int k = 0; k < NPL; ++k) {
                const int i = lane + 32 * k;
                const bool start = ((inact >> i) & 1ull) && (i == 0 || !((inact >> (i - 1)) & 1ull));
                if (start) {
                    const unsigned long long rest = ~inact >> i;
                    const int L = rest ? __ffsll((long long)rest) - 1 : 64 - i;
                    tri_solve_run(S, i, L);
                }
            }
            __syncwarp();
            double ss[NPL];
#pragma unroll
            for (int k = 0; k < NPL; ++k) {  // step_finish_kernel
                const int i = lane + 32 * k;
                ss[k] = 0.0;
                out[k] = base[k];
                if (i < n) {
                    double dx = act[k] ? __dmul_rn(dtb, Fv[k]) : S.R[i];
                    const double xj = base[k];
                    double v = __dsub_rn(xj, dx);
                    if (v < lb[k]) { v = lb[k]; dx = __dsub_rn(xj, lb[k]); }
                    if (v > ub[k]) { v = ub[k]; dx = __dsub_rn(xj, ub[k]); }
                    out[k] = v;
                    ss[k] = fma(dx, dx, 0.0);
                }
            }
            __syncwarp();
            return sqrt(tree_sum(ss));
        };
        // Problem callbacks at `pt` (into S.xs) and, optionally, |F_unscaled(pt)| w.r.t. the current iterate with the
        // active set recomputed at pt (residual_kernel, unscaled, mode 0).
        auto eval_point = [&](const double (&pt)[NPL], double (&gp)[NPL], double& op, bool want_norm) -> double {
#pragma unroll
            for (int k = 0; k < NPL; ++k) S.xs[lane + 32 * k] = pt[k];
            __syncwarp();
            rosen_eval(S, n, lane, gp, op);
            if (!want_norm) return 0.0;
            double ss[NPL];
#pragma unroll
            for (int k = 0; k < NPL; ++k) {
                ss[k] = 0.0;
                if (lane + 32 * k < n) {
                    const double p = __dsub_rn(x[k], __dmul_rn(dtb, gp[k]));
                    const bool a2 = (p < lb[k] - GF_ACTIVE_SLACK) || (p > ub[k] + GF_ACTIVE_SLACK);
                    const double proj = a2 ? fmin(fmax(p, lb[k]), ub[k]) : p;
                    const double rx = __dsub_rn(pt[k], proj);
                    ss[k] = fma(rx, rx, 0.0);
                }
            }
            return sqrt(tree_sum(ss));
        };

        if (failed) {  // step_control.py:102-104
            phase = GF_PHASE_FAILED;
            lamb_next = 2.0 * lamb;
        } else {
            diff1 = newton_step(x, F, xm);
            const double mid_norm = eval_point(xm, gm, objm, true);
            if (mid_norm <= prm.newton_tol) {          // distance_ratio_control.py:34-39
                phase = GF_PHASE_ACCEPT_MID;
                lamb_next = fmax(lamb * prm.lamb_red, prm.lamb_min);
            } else if (diff1 == 0.0) {                 // :41-44
                phase = GF_PHASE_ACCEPT_MID;
                lamb_next = lamb;
            } else {
                // ---- second step from mid, active set and factor frozen (residual_kernel, scaled, mode 1) ----
                double F2[NPL];
#pragma unroll
                for (int k = 0; k < NPL; ++k) {
                    const double p = __dsub_rn(__dmul_rn(lamb, x[k]), gm[k]);
                    const double lo = __dmul_rn(lamb, lb[k]), hi = __dmul_rn(lamb, ub[k]);
                    const double proj = act[k] ? fmin(fmax(p, lo), hi) : p;
                    F2[k] = __dsub_rn(__dmul_rn(lamb, xm[k]), proj);
                }
                diff2 = newton_step(xm, F2, xf);
                if (diff2 == 0.0) {                    // :50-53
                    phase = GF_PHASE_ACCEPT_FINAL;
                    lamb_next = lamb;
                } else {
                    const double theta = diff2 / diff1;
                    if (theta <= prm.theta_max) {      // :57-63, controller.py:44-77
                        const double err = prm.log_theta_ref - log(theta);
                        const double es = err_sum + err;
                        err_sum = es;
                        const double mod = exp(prm.K_P * err + prm.K_I * es);
                        lamb_next = fmax(prm.lamb_min, lamb / mod);
                        phase = GF_PHASE_ACCEPT_FINAL;
                    } else {
                        lamb_next = lamb * prm.lamb_inc;  // :65
                        phase = GF_PHASE_REJECT;
                    }
                }
                if (phase == GF_PHASE_ACCEPT_FINAL) eval_point(xf, gf, objf, false);
            }
        }
        nsteps += (phase >= 2 && phase <= 4 ? 1 : 0) + (phase == 3 || phase == 4 ? 1 : 0);
        // ---- commit (commit_kernel) ----
        if (lamb_next >= prm.lamb_max) {  // solver.py:323-326
            status = GF_STATUS_LAMB_MAX;
            lamb_cur = lamb_next;
            break;
        }
        if (phase == GF_PHASE_ACCEPT_MID) {
#pragma unroll
            for (int k = 0; k < NPL; ++k) { x[k] = xm[k]; g[k] = gm[k]; }
            obj = objm;
            ++accepted;
        } else if (phase == GF_PHASE_ACCEPT_FINAL) {
#pragma unroll
            for (int k = 0; k < NPL; ++k) { x[k] = xf[k]; g[k] = gf[k]; }
            obj = objf;
            ++accepted;
        }
        lamb_cur = lamb_next;
        ++iters;
        // the next iteration reads the neighbours of the current iterate from S.xs
#pragma unroll
        for (int k = 0; k < NPL; ++k) S.xs[lane + 32 * k] = x[k];
        __syncwarp();
    }

#pragma unroll
    for (int k = 0; k < NPL; ++k) {
        const int i = lane + 32 * k;
        if (i < n) {
            x_all[(size_t)b * n + i] = x[k];
            grad_all[(size_t)b * n + i] = g[k];
            if (active_all != nullptr) active_all[(size_t)b * n + i] = (amask >> i) & 1ull;
        }
    }
    if (lane == 0) {
        obj_all[b] = obj;
        lamb_all[b] = lamb_cur;
        errsum_all[b] = err_sum;
        status_all[b] = status;
        iters_all[b] = iters;
        accepted_all[b] = accepted;
        nsteps_all[b] = nsteps;
        totres_all[b] = total_res;
    }
}

}  // namespace

extern "C" int gf_rosen_fused_solve(int B, int n, const double* a, const double* b, const double* lb, const double* ub,
                                    double* x, double* grad, double* obj, double* lamb, double* err_sum,
                                    int32_t* status, int32_t* iters, int32_t* accepted, int32_t* newton_steps,
                                    double* total_res, uint8_t* active, const double* params_host, int iteration_limit,
                                    int max_outer, int fresh, const int32_t* work, const int32_t* nwork_dev, int nwork,
                                    void* stream) {
    if (B <= 0 || n < 2 || !a || !b || !lb || !ub || !x || !grad || !obj || !lamb || !err_sum || !status || !iters ||
        !accepted || !newton_steps || !total_res || !params_host || max_outer < 1 || nwork < 0)
        return GF_ERR_ARG;
    if (n > FN) return GF_ERR_UNSUPPORTED;
    if (nwork == 0) return GF_OK;
    FusedPrm p;
    p.opt_tol = params_host[0];
    p.active_tol = params_host[1];
    p.obj_lower_limit = params_host[2];
    p.newton_tol = params_host[3];
    p.lamb_red = params_host[4];
    p.lamb_min = params_host[5];
    p.lamb_max = params_host[6];
    p.lamb_inc = params_host[7];
    p.theta_max = params_host[8];
    p.log_theta_ref = params_host[9];
    p.K_P = params_host[10];
    p.K_I = params_host[11];
    p.iteration_limit = iteration_limit;
    fused_rosen_kernel<<<(nwork + FW - 1) / FW, FW * 32, 0, (cudaStream_t)stream>>>(
        n, a, b, lb, ub, x, grad, obj, lamb, err_sum, status, iters, accepted, newton_steps, total_res, active, p,
        max_outer, fresh, GfWork{work, nwork_dev}, nwork);
    return gf_launch_status();
}
