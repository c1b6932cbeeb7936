/* libgradflow_b200 -- C ABI of the B200-native batched Newton/KKT path of pygradflow.
 *
 * Drop-in boundary.  The reference is pure Python; its plug-in points for this path are
 *   - Params.step_solver (pygradflow/params.py:234), called as step_solver(problem, params, iterate, dt, rho)
 *     by pygradflow/step/solver/__init__.py:18-19 and expected to return a StepSolver
 *     (pygradflow/step/solver/step_solver.py:66-130);
 *   - StepSolver.linear_solver(mat) (step_solver.py:94-98) returning a LinearSolver
 *     (pygradflow/linear_solver/linear_solver.py:18-31).
 * A maintainer binds this library with ctypes; pygradflow_b200/native.py is that
 * binding.  Each entry point below names the reference code it replaces.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host; buffers are owned by the caller
 *     (torch tensors); the library allocates nothing except inside the *_host entry points' workspace,
 *     which the caller also provides;
 *   - all floating point is IEEE binary64, matrices row-major and contiguous, batch index outermost:
 *     x[B,n] y[B,m] grad[B,n] cons[B,m] J[B,m,n] H[B,n,n] (symmetric) K[B,ld,ld] rhs[B,ld];
 *   - per-instance scalars are arrays: dt[B] (= 1/lambda, the reference passes dt and recomputes
 *     lambda = 1/dt: implicit_func.py:212), rho[B];
 *   - work / nwork_dev / nwork: optional work list.  The batch dimension of the grid has `nwork` CTAs;
 *     CTA w handles instance work[w] (or w when work == NULL) and exits if nwork_dev != NULL and
 *     w >= *nwork_dev.  This is how finished / rejected instances are skipped without a host sync;
 *   - `stream` is a cudaStream_t passed as void*; every call is asynchronous on it;
 *   - return value: 0 ok, GF_ERR_ARG bad argument, GF_ERR_UNSUPPORTED shape not supported,
 *     1000 + cudaError_t on a launch error.  Numerical failures are per instance in info[B].
 */
#ifndef GRADFLOW_B200_H
#define GRADFLOW_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* status words: SolverStatus of pygradflow/status.py:4-31 (enum auto() values); 0 = still running */
#define GF_STATUS_RUNNING 0
#define GF_STATUS_OPTIMAL 1
#define GF_STATUS_ITERATION_LIMIT 2
#define GF_STATUS_TIME_LIMIT 3
#define GF_STATUS_UNBOUNDED 4
#define GF_STATUS_LOCALLY_INFEASIBLE 5
/* solver.py:323-326 raises when lambda >= lamb_max; a batch records it per instance instead */
#define GF_STATUS_LAMB_MAX 6
/* newton.py:294 raises when the Armijo search is exhausted */
#define GF_STATUS_LINE_SEARCH_FAILED 7

/* phase of an instance inside one outer iteration (DistanceRatioController.step) */
#define GF_PHASE_IDLE 0
#define GF_PHASE_SECOND 1
#define GF_PHASE_ACCEPT_MID 2
#define GF_PHASE_ACCEPT_FINAL 3
#define GF_PHASE_REJECT 4
#define GF_PHASE_FAILED 5

/* info[B] of the factorisations: 0 ok, k > 0 zero/non-finite pivot at column k, -1 non-finite matrix,
 * -2 pivot signs are not those of a quasi-definite matrix (unpivoted LDL' not trusted; refactor with LU) */
#define GF_INFO_NONFINITE (-1)
#define GF_INFO_NOT_QUASIDEFINITE (-2)

int gf_version(void);

/* ---- problem-family evaluators (Problem callbacks: pygradflow/problem.py:112-192, iterate.py:59-76) ---- */

/* QP family (tests/pygradflow/qp.py:17-30): grad = H x + g, cons = A x + b, obj = x'Hx/2 + g'x. */
int gf_qp_eval(int B, int n, int m, const double* H, const double* A, const double* g, const double* b,
               const double* x, double* grad, double* cons, double* obj, const int32_t* work,
               const int32_t* nwork_dev, int nwork, void* stream);

/* chained Rosenbrock (n = 2 is tests/pygradflow/rosenbrock.py:15-46): gradient / objective, and the three
 * diagonals of the dense Hessian (H must be zero elsewhere). */
int gf_rosen_eval(int B, int n, const double* a, const double* b, const double* x, double* grad, double* obj,
                  const int32_t* work, const int32_t* nwork_dev, int nwork, void* stream);
int gf_rosen_hess(int B, int n, const double* a, const double* b, const double* x, double* H,
                  const int32_t* work, const int32_t* nwork_dev, int nwork, void* stream);

/* discretised optimal-control family (cfg4 of the baseline; synthetic, the reference ships no OCP -- the callbacks
 * it replaces are Problem.obj / obj_grad / cons / cons_jac / lag_hess, problem.py:112-192): variables
 * z = (x_1, u_0, ..., x_S, u_{S-1}), c_j = x_{j+1} - x_j - h (A_j x_j + B_j u_j + 0.1 sin x_j), x_0 = xinit,
 * cost 1/2 sum (x'Qx + u'Ru) with diagonal Q [B,S,nx], R [B,S,nu]; A [B,S,nx,nx], Bm [B,S,nx,nu].
 * gf_ocp_jac / gf_ocp_hess write only the non-zero blocks / the diagonal of dense J [B,m,n] / H [B,n,n]
 * (zero elsewhere); c1 = 0.1 h. */
int gf_ocp_eval(int B, int S, int nx, int nu, double h, const double* A, const double* Bm, const double* Q,
                const double* R, const double* xinit, const double* z, double* grad, double* cons, double* obj,
                const int32_t* work, const int32_t* nwork_dev, int nwork, void* stream);
int gf_ocp_jac(int B, int S, int nx, int nu, double h, const double* A, const double* Bm, const double* z, double* J,
               const int32_t* work, const int32_t* nwork_dev, int nwork, void* stream);
int gf_ocp_hess(int B, int S, int nx, int nu, double c1, const double* Q, const double* R, const double* z,
                const double* y, double* H, const int32_t* work, const int32_t* nwork_dev, int nwork, void* stream);

/* Iterate.aug_lag_deriv_x (iterate.py:91-94): dL = grad + J'(rho c + y); optionally J'y (iterate.py:138,171)
 * and J'c (iterate.py:125) from the same pass over J.  dL / jty / jtc may be NULL. */
int gf_aug_lag_grad(int B, int n, int m, const double* J, const double* grad, const double* cons, const double* y,
                    const double* rho, double* dL, double* jty, double* jtc, const int32_t* work,
                    const int32_t* nwork_dev, int nwork, void* stream);

/* ---- residual, active set (implicit_func.py) ---- */

/* StepFunc.value_at / compute_active_set (implicit_func.py:21-60,150-161,219-252).
 * scaled != 0: ScaledImplicitFunc, else ImplicitFunc.  active_mode 0: recompute the active set from
 * p (stored to `active` when non-NULL), 1: use `active`.  F[B,n+m] and nrm[B] (= ||F||_2) may be NULL.
 * dL / cons are evaluated at (x, y); (x0, y0) is the iterate the implicit-Euler step started from. */
int gf_residual(int B, int n, int m, const double* x, const double* y, const double* x0, const double* y0,
                const double* dL, const double* cons, const double* lb, const double* ub, const double* dt,
                int scaled, int active_mode, uint8_t* active, double* F, double* nrm, const int32_t* work,
                const int32_t* nwork_dev, int nwork, void* stream);
/* the same with the tau-variant of the active-set decision: tau[B] from ActiveSetType Explicit / Smallest / Largest
 * (newton_control.py:60-88); the active set is decided on f_x x + f_x0 x0 - f_d dL (implicit_func.py:237-244),
 * the residual is unchanged.  tau == NULL is gf_residual. */
int gf_residual_tau(int B, int n, int m, const double* x, const double* y, const double* x0, const double* y0,
                    const double* dL, const double* cons, const double* lb, const double* ub, const double* dt,
                    const double* tau, int scaled, int active_mode, uint8_t* active, double* F, double* nrm,
                    const int32_t* work, const int32_t* nwork_dev, int nwork, void* stream);

/* np.where(~active)[0] / np.where(active)[0] (scaled_step_solver.py:51-52): perm[b] = inactive indices
 * ascending followed by active indices ascending, nI[b] = #inactive, Nvec[b] = nI[b] + m (may be NULL). */
int gf_index_sets(int B, int n, int m, const uint8_t* active, int32_t* perm, int32_t* nI, int32_t* Nvec,
                  const int32_t* work, const int32_t* nwork_dev, int nwork, void* stream);

/* ---- KKT assembly (step/solver/symmetric_step_solver.py) ---- */

/* compute_hess_jac + _compute_deriv (symmetric_step_solver.py:27-39,49-77):
 * K = [[H[I,I] + lamb I, J[:,I]'], [J[:,I], -lamb/(1+lamb rho) I]], order N = nI + m, top-left of K[b];
 * rows/cols N..roundup(N,pad)-1 are set to identity.  lower_only != 0 writes only the lower triangle. */
int gf_kkt_assemble(int B, int n, int m, int ld, int pad, int lower_only, const double* H, const double* J,
                    const int32_t* perm, const int32_t* nI, const double* dt, const double* rho, double* K,
                    const int32_t* work, const int32_t* nwork_dev, int nwork, void* stream);

/* initial_rhs + compute_rhs (scaled_step_solver.py:38-60,91-97, symmetric_step_solver.py:79-94):
 * rhs = [F_x[I] - H[I,A] (dt F_x[A]);  fact F_y - J[:,A] (dt F_x[A])],  fact = 1/(1 + lamb rho). */
int gf_kkt_rhs(int B, int n, int m, int ld, const double* H, const double* J, const int32_t* perm, const int32_t* nI,
               const double* F, const double* dt, const double* rho, double* rhs, const int32_t* work,
               const int32_t* nwork_dev, int nwork, void* stream);

/* The unsymmetric full-order formulations of the scaled step (step/solver/__init__.py:12-31): K[b] (ld >= n + m) gets
 * the matrix of AsymmetricStepSolver.compute_deriv (asymmetric_step_solver.py:77-104, unit rows for the active
 * variables :37-75) or of ExtendedStepSolver._compute_deriv (extended_step_solver.py:39-83); gf_kkt_rhs_full the
 * matching right-hand side (asymmetric_step_solver.py:106-123 / extended_step_solver.py:93).  The solution is
 * (dx, sy) in the natural order for both: gf_step_finish with the identity permutation and nI = n finishes it. */
#define GF_FORM_SYMMETRIC 0
#define GF_FORM_ASYMMETRIC 1
#define GF_FORM_EXTENDED 2
#define GF_FORM_STANDARD 3 /* unscaled implicit function (standard_step_solver.py:15-92, implicit_func.py:163-199): H must
                              be H_rho = lag_hess(x, y + rho c) + rho J'J, rhs = F unscaled, solution = (dx, dy) */
#define GF_FORM_SCALED_DERIV 4 /* gf_kkt_assemble_full only: F' of the scaled implicit function itself,
                                 [[lamb I + P_I H_rho, P_I J'], [-J, lamb I]] (ScaledImplicitFunc.deriv, implicit_func.py:254-294;
                                 what GlobalizedNewtonMethod reads through StepFunc.deriv_at, newton.py:262) */
int gf_kkt_assemble_full(int B, int n, int m, int ld, int form, const double* H, const double* J, const int32_t* perm,
                         const int32_t* nI, const uint8_t* active, const double* dt, const double* rho, double* K,
                         const int32_t* work, const int32_t* nwork_dev, int nwork, void* stream);
int gf_kkt_rhs_full(int B, int n, int m, int ld, int form, const int32_t* perm, const int32_t* nI,
                    const uint8_t* active, const double* F, const double* dt, const double* rho, double* rhs,
                    const int32_t* work, const int32_t* nwork_dev, int nwork, void* stream);

/* ---- linear solver (linear_solver/lu_solver.py:9-21; LinearSolver contract linear_solver.py:18-31) ---- */

/* LUSolver.__init__: in-place LU with partial pivoting of the order-Nvec[b] (or Nmax when Nvec == NULL)
 * matrix in K[b]; piv[B,ld] = the interchange sequence (LAPACK getrf's, on the transposed view); info[B] = 0,
 * k > 0 (zero pivot -- |pivot| < 1e-290 -- in column k, 1-based: the reference's "exactly singular" RuntimeError,
 * lu_solver.py:15-17) or
 * -1 (non-finite).  The packed factors are only meaningful to gf_lu_solve: the interchanges of a 32-column block are
 * not applied to the blocks of L left of it (the substitution applies them block by block). */
int gf_lu_factor(int B, int ld, int Nmax, const int32_t* Nvec, double* K, int32_t* piv, int32_t* info,
                 const int32_t* work, const int32_t* nwork_dev, int nwork, void* stream);
/* LUSolver.solve(rhs, trans): rhs[B,ldr] is overwritten by the solution of K x = rhs (trans = 0) or
 * K' x = rhs (trans != 0; only cond_estimate.py:82 uses it). */
int gf_lu_solve(int B, int ld, int Nmax, const int32_t* Nvec, const double* K, const int32_t* piv, double* rhs,
                int ldr, int trans, const int32_t* work, const int32_t* nwork_dev, int nwork, void* stream);

/* Symmetric factorisation with inertia (contract of ma57_solver.py:76-79, mumps_solver.py:81-82,
 * ssids_solver.py:22-23, cholesky_solver.py:21-22): unpivoted L D L' of the lower triangle, ld % 64 == 0,
 * matrix padded with identity to a multiple of 64 (gf_kkt_assemble pad = 64).  dvec[B,ld] receives D,
 * nneg[B] the number of negative pivots (= num_neg_eigvals).  npos_expected (may be NULL): when given, a
 * pivot whose sign differs from (+ for the first npos_expected[b], - after) sets info = -2. */
int gf_ldlt_factor(int B, int ld, int Nmax, const int32_t* Nvec, double* K, double* dvec, int32_t* info,
                   int32_t* nneg, const int32_t* npos_expected, const int32_t* work, const int32_t* nwork_dev,
                   int nwork, void* stream);
/* gf_kkt_assemble (lower triangle, padded to 64) fused into the factorisation: every tile of K is gathered from H, J
 * and the index sets at its first touch instead of being written and read back once (symmetric_step_solver.py:27-39,
 * 49-77 + lu_solver.py:9-17 in one call).  Same result as gf_kkt_assemble(pad 64, lower only) + gf_ldlt_factor with
 * npos_expected = nI, bit for bit. */
int gf_kkt_ldlt_factor(int B, int n, int m, int ld, const double* H, const double* J, const int32_t* perm,
                       const int32_t* nI, const double* dt, const double* rho, const int32_t* Nvec, double* K,
                       double* dvec, int32_t* info, int32_t* nneg, const int32_t* work, const int32_t* nwork_dev,
                       int nwork, void* stream);
int gf_ldlt_solve(int B, int ld, int Nmax, const int32_t* Nvec, const double* K, double* rhs, int ldr,
                  const int32_t* work, const int32_t* nwork_dev, int nwork, void* stream);

/* ---- step finish (scaled_step_solver.py:99-107, symmetric_step_solver.py:113-121, step_solver.py:16-63) ---- */

/* dx[I] = sol[:nI], dx[A] = dt F_x[A], dy = fact (sol[nI:] - rho F_y); xn = clip(xbase - dx) with the dx
 * fix-up, yn = ybase - dy, diff = ||(dx, dy)||_2.  dx / dy outputs may be NULL. */
int gf_step_finish(int B, int n, int m, int ld, const double* xbase, const double* ybase, const double* sol,
                   const int32_t* perm, const int32_t* nI, const double* F, const double* dt, const double* rho,
                   const double* lb, const double* ub, double* xn, double* yn, double* dx, double* dy, double* diff,
                   const int32_t* work, const int32_t* nwork_dev, int nwork, void* stream);

/* ---- callers restated per instance (needed for identical iteration counts / status) ---- */

/* Solver._check_terminate (solver.py:180-205) with Iterate.total_res / locally_infeasible
 * (iterate.py:115-181) and ActiveSet (active_set.py:4-29).  iteration_limit < 0: none.  Only instances with
 * status == 0 are examined. */
int gf_check_terminate(int B, int n, int m, const double* x, const double* grad, const double* cons,
                       const double* jty, const double* jtc, const double* obj, const double* lb, const double* ub,
                       double opt_tol, double active_tol, double local_infeas_tol, double obj_lower_limit,
                       int iteration_limit, const int32_t* iters, int32_t* status, double* total_res,
                       const int32_t* work, const int32_t* nwork_dev, int nwork, void* stream);

/* DistanceRatioController.step (distance_ratio_control.py:27-44) after the first Newton step, plus the
 * solver-failure path of StepController.compute_step (step_control.py:80-83,102-104) when info != 0. */
int gf_dr_first(int B, const int32_t* status, const int32_t* info, const double* dt, const double* mid_norm,
                const double* diff1, double newton_tol, double lamb_red, double lamb_min, int32_t* phase,
                double* lamb_next, void* stream);
/* ... after the second step (distance_ratio_control.py:46-78) with LogController (controller.py:44-77). */
int gf_dr_second(int B, const double* dt, const double* diff1, const double* diff2, double theta_max,
                 double log_theta_ref, double K_P, double K_I, double lamb_min, double lamb_inc, double* err_sum,
                 int32_t* phase, double* lamb_next, double* theta, void* stream);
/* Newton steps of one DistanceRatio outer iteration (the driver's display counter): adds, over all instances, one step for
 * phase 2..4 and a second one for phase 3, 4 to nsteps (device int64). */
int gf_count_newton_steps(int B, const int32_t* phase, int64_t* nsteps, void* stream);
/* ResiduumRatioController.step (residuum_ratio_control.py:18-63; fixed != 0: FixedStepSizeController.step,
 * fixed_control.py:12-19) after their single Newton step; nsteps (device int64) counts the Newton steps taken. */
int gf_single_control(int B, int fixed, const int32_t* status, const int32_t* info, const double* dt,
                      const double* mid_norm, const double* orig_norm, double newton_tol, double theta_max,
                      double log_theta_ref, double K_P, double K_I, double lamb_red, double lamb_min, double lamb_inc,
                      double lamb_init, double* err_sum, int32_t* phase, double* lamb_next, double* theta,
                      int64_t* nsteps, void* stream);
/* ExactController.step (exact_control.py:16-66), one call per stage of its Newton loop over the instances with
 * live[b] != 0.  mode 0: set-up after the first step (info != 0 => failed); mode 1: verdict on Newton iterate `it`
 * (residual norm val[B]; even iterates are committed from the mid buffers, odd from the fin buffers; last != 0: the
 * tenth iterate); mode 2: a failed refactorisation ends an instance's loop. */
int gf_exact_control(int B, int mode, int it, int last, const int32_t* status, const int32_t* info, const double* dt,
                     const double* val, const double* orig_norm, double newton_tol, double rate_bound, double* curr,
                     int32_t* live, int32_t* phase, double* lamb_next, int64_t* nsteps, void* stream);
/* End of the outer iteration (solver.py:318-378): lamb_max guard, penalty update (dual_norm_update = 0: constant,
 * 1: DualNormUpdate penalty.py:46-74, 2: DualEquilibration penalty.py:77-113), iterate <- accepted Newton iterate,
 * counters. */
int gf_commit(int B, int n, int m, const int32_t* phase, const double* lamb_next, double lamb_max,
              int dual_norm_update, const double* xm, const double* ym, const double* gm, const double* cm,
              const double* om, const double* xf, const double* yf, const double* gf, const double* cf,
              const double* of, double* x, double* y, double* grad, double* cons, double* obj, double* lamb,
              double* rho, int32_t* iters, int32_t* accepted, int32_t* status, void* stream);

/* ---- globalized Newton (newton.py:242-304): merit gradient, trial point, fused residual-norm + Armijo ---- */

/* res = 1/2 |F|^2 and inner = (F'^T F) . (dx, dy) (newton.py:254,262-271) without assembling F'
 * (implicit_func.py:254-294): H is the Hessian of the Lagrangian at multiplier y + rho c, F the scaled residual at
 * the current iterate, `active` its active set. */
int gf_merit_grad(int B, int n, int m, const double* H, const double* J, const double* F, const uint8_t* active,
                  const double* dt, const double* rho, const double* dx, const double* dy, double* res, double* inner,
                  const int32_t* work, const int32_t* nwork_dev, int nwork, void* stream);
/* Start of the Armijo search for the instances of the work list (newton.py:256-257,273): state = 1 (full step) where
 * res <= newton_tol, else 0 (searching); alpha = 1, trials = 0.  Instances outside the list keep their state. */
int gf_ls_begin(int B, const double* res, double newton_tol, int32_t* state, double* alpha, int32_t* trials,
                const int32_t* work, const int32_t* nwork_dev, int nwork, void* stream);
/* xt = x - alpha dx, yt = y - alpha dy (newton.py:276-278) */
int gf_ls_trial(int B, int n, int m, const double* x, const double* y, const double* dx, const double* dy,
                const double* alpha, double* xt, double* yt, const int32_t* work, const int32_t* nwork_dev, int nwork,
                void* stream);
/* Scaled residual at the trial point, 1/2 |F|^2 (warp-shuffle reduction) and the Armijo test / step halving of
 * newton.py:280-290 in one kernel.  state[B]: 0 searching, 1 accepted at alpha[B], 2 exhausted after max_trials
 * (newton.py:294 raises). */
int gf_armijo_residual(int B, int n, int m, const double* xt, const double* yt, const double* x0, const double* y0,
                       const double* dL, const double* cons, const double* lb, const double* ub, const double* dt,
                       const double* res, const double* inner, double newton_tol, int max_trials, double* alpha,
                       int32_t* trials, int32_t* state, double* next_res, const int32_t* work,
                       const int32_t* nwork_dev, int nwork, void* stream);

/* ---- banded L D L' for families whose KKT matrix is banded under a given ordering (cfg4: discretised optimal
 * control).  Replaces lu_solver.py:9-21 on the matrix of symmetric_step_solver.py:49-77, assembled as the FULL
 * (n + m) system in the family's order with identity rows / columns for the active variables, so order and band do
 * not depend on the active set.  order[t] = full KKT index (variables 0..n-1, constraints n..n+m-1) at band
 * position t, pos = its inverse; Kband [B, n+m, bw+1] with Kband[b][t][d] = K(t, t-d); bw + 1 even and <= 64.
 * gf_band_permute moves a vector between the reduced standard order (inactive variables in perm order, then
 * constraints; leading dimension ld) and the band order. */
int gf_band_assemble(int B, int n, int m, int bw, const double* H, const double* J, const uint8_t* active,
                     const int32_t* order, const double* dt, const double* rho, double* Kband, const int32_t* work,
                     const int32_t* nwork_dev, int nwork, void* stream);
int gf_band_factor(int B, int N, int bw, double* Kband, int32_t* info, int32_t* nneg, const int32_t* work,
                   const int32_t* nwork_dev, int nwork, void* stream);
int gf_band_solve(int B, int N, int bw, const double* Kband, double* v, const int32_t* work, const int32_t* nwork_dev,
                  int nwork, void* stream);
int gf_band_permute(int B, int n, int m, int ld, const int32_t* perm, const int32_t* nI, const int32_t* pos,
                    double* stdv, double* bandv, int to_band, const int32_t* work, const int32_t* nwork_dev, int nwork,
                    void* stream);

/* ---- stage-structured KKT systems (cfg4: discretised optimal control, n = S (nx + nu), m = S nx; nx == 8) ------------
 * Compact layouts: Jc [B, S, nx + w, nx], w = nx + nu, column-major inside a stage -- Jc[b][j][c][r] = d c_{j,r} / d (column
 * c), columns 0..nx-1 = x_j (variables (j-1) w + c, zero block for j = 0), columns nx..nx+w-1 = z_j (variables j w + c - nx);
 * Hd [B, n] the DIAGONAL Hessian of the Lagrangian.  gf_ocp_jac_banded / gf_ocp_hess_diag are the compact twins of
 * gf_ocp_jac / gf_ocp_hess (diag_only != 0: Jc already holds the Jacobian of an earlier point, only the entries that
 * depend on z are rewritten); gf_stage_aug_lag_grad is gf_aug_lag_grad (iterate.py:91-94,125,138,171) on the compact
 * Jacobian.
 * gf_stage_kkt_factor replaces symmetric_step_solver.py:27-77 + lu_solver.py:9-17 for such families: it forms the Schur
 * complement on the multipliers M = delta I + J_I (H_II + lamb I)^-1 J_I' (SPD, block tridiagonal with S blocks of
 * nx x nx; the same "primal variables first" elimination order as the dense LDL') and factorises it by block cyclic
 * reduction -- Tinv / P / Q [B, S, nx*nx] receive, for every block i, the inverse pivot block and Tinv K[i, i -/+ s] at the
 * level s where it is eliminated; nu must be a multiple of 4, <= 16; info = 0, -2 (some H_ii + lamb <= 0: K not
 * quasi-definite) or k > 0 (pivot breakdown).
 * gf_stage_kkt_solve does scaled_step_solver.py:38-60,91-97 + symmetric_step_solver.py:79-121 + lu_solver.py:19-21 for
 * the scaled residual F: sol [B, ldsol] = (dx in the natural order with dx_A = dt F_x[A], sy); gf_step_finish with the
 * identity permutation and nI = n finishes the step. */
int gf_ocp_jac_banded(int B, int S, int nx, int nu, double h, const double* A, const double* Bm, const double* z,
                      double* Jc, int diag_only, const int32_t* work, const int32_t* nwork_dev, int nwork, void* stream);
int gf_ocp_hess_diag(int B, int S, int nx, int nu, double c1, const double* Q, const double* R, const double* z,
                     const double* y, double* Hd, const int32_t* work, const int32_t* nwork_dev, int nwork, void* stream);
int gf_stage_aug_lag_grad(int B, int S, int nx, int nu, const double* Jc, const double* grad, const double* cons,
                          const double* y, const double* rho, double* dL, double* jty, double* jtc, const int32_t* work,
                          const int32_t* nwork_dev, int nwork, void* stream);
int gf_stage_kkt_factor(int B, int S, int nx, int nu, const double* Jc, const double* Hd, const uint8_t* active,
                        const double* dt, const double* rho, double* Tinv, double* P, double* Q, int32_t* info,
                        int32_t* nneg, const int32_t* work, const int32_t* nwork_dev, int nwork, void* stream);
int gf_stage_kkt_solve(int B, int S, int nx, int nu, const double* Jc, const double* Hd, const uint8_t* active,
                       const double* F, const double* dt, const double* rho, const double* Tinv, const double* P,
                       const double* Q, double* sol, int ldsol, const int32_t* work, const int32_t* nwork_dev, int nwork,
                       void* stream);

/* ---- host-buffer entry (the reference hands its step solver host arrays: scaled_step_solver.py:76-79) ---- */

/* ScaledProblem (scale.py:153-231): out[b][r][c] = ldexp(in[b][r][c], sr rw[b][r] + sc cw[b][c] + so ow[b]) for
 * the instances of the work list; rw [B, rows], cw [B, cols], ow [B] are the integer weights of the reference's
 * Scaling (var_weights / cons_weights / obj_weight), any of them may be NULL.  Exact (powers of two). */
int gf_ldexp(int B, int rows, int cols, const double* in, const int32_t* rw, int sr, const int32_t* cw, int sc,
             const int32_t* ow, int so, double* out, const int32_t* work, const int32_t* nwork_dev, int nwork,
             void* stream);

/* H [cnt, n, n] is symmetric (problem.py:174-192): copy only its lower block triangle (row blocks of `blk` rows,
 * columns up to the end of the diagonal block) from pinned host memory with one strided copy per row block, and
 * rebuild the blocks above the diagonal on the device (blk a multiple of 32). */
int gf_h2d_sym_lower(double* dst, const double* src_host, int cnt, int n, int blk, void* stream);
int gf_symmetrize_lower(double* H, int cnt, int n, int blk, void* stream);

/* Hessian block of the Standard formulation: out[B,n,n] = H + rho[b] J'J (iterate.py:103-110 aug_lag_deriv_xx(rho), used by
 * standard_step_solver.py:50-53); H = lag_hess(x, y + rho c) comes from the family.  Batched rank-m update on the FP64
 * tensor pipe.  nwork <= 65535. */
int gf_hess_rho(int B, int n, int m, const double* H, const double* J, const double* rho, double* out,
                const int32_t* work, const int32_t* nwork_dev, int nwork, void* stream);

/* ---- iterative LinearSolvers (linear_solver/gmres_solver.py:7-35, linear_solver/minres_solver.py:6-24) ----
 * One CTA per instance runs the whole iteration of scipy.sparse.linalg.gmres / minres on the dense K[b] (order Nvec[b] or
 * Nmax, row-major, ld >= Nmax) and overwrites rhs[B,ldr] with the solution.  info[b] = 0 on convergence, otherwise the
 * iteration limit (the value scipy returns and the wrappers turn into LinearSolverError: gmres_solver.py:32-33,
 * minres_solver.py:21-22); iters[b] (may be NULL) = matrix-vector products scipy's routine performs.
 * scratch: gf_krylov_scratch_rows(method, restart) * ld doubles per instance (indexed by instance, not by work slot).
 * GMRES: restart = 20 and rtol = 1e-5 are scipy's defaults, atol = 1e-8 and maxiter = N the wrapper's arguments;
 * trans != 0 solves K' x = rhs (LinearSolver.solve(trans=True)).  The start vector (LinearSolver.solve(initial_sol=...),
 * a zero-argument callable in the reference) is x0[B,ldr], or -- x0 == NULL, x0_mask[B,nmask] != NULL -- the vector
 * AsymmetricStepSolver.initial_sol builds (asymmetric_step_solver.py:125-138): rhs[i] where x0_mask[i] is set (i < nmask),
 * zero elsewhere; both NULL = no start vector.  With a start vector the wrapper's early return applies (:22-25).
 * MINRES: shift 0, rtol = 1e-5, maxiter = 5 N, the matrix must be symmetric (minres_solver.py:9). */
#define GF_KRYLOV_GMRES 0
#define GF_KRYLOV_MINRES 1
int gf_krylov_scratch_rows(int method, int restart);
int gf_gmres_solve(int B, int ld, int Nmax, const int32_t* Nvec, const double* K, double* rhs, int ldr, const double* x0,
                   const uint8_t* x0_mask, int nmask, int trans, int restart, double rtol, double atol, double* scratch,
                   int32_t* info, int32_t* iters, const int32_t* work, const int32_t* nwork_dev, int nwork,
                   void* stream);
int gf_minres_solve(int B, int ld, int Nmax, const int32_t* Nvec, const double* K, double* rhs, int ldr, const double* x0,
                    double rtol, double* scratch, int32_t* info, int32_t* iters, const int32_t* work,
                    const int32_t* nwork_dev, int nwork, void* stream);

/* ---- penalty strategies that look at the candidate iterate (penalty.py:115-255; veto of solver.py:357-378) ----
 * gf_pareto_update: ParetoDecrease.update (penalty.py:136-168) for every instance of the work list whose last step was
 * accepted (phase 2 / 3, status running), on the committed iterate: grad = obj_grad, jty = J'y, jtc = J'c (the products
 * gf_aug_lag_grad delivers); rho[b] <- max(min(10 rho, bound), rho).  Never rejects.
 * gf_filter_update: kind 0 = ObjectivePenaltyFilter (entry = (obj, |c|_inf)), kind 1 = LagrangianPenaltyFilter (entry =
 * (|dL|^2 + |c|^2, |c|_2), dL = aug_lag_deriv_x of the candidate at the strategy's rho_pen[b], evaluated by the caller).
 * The candidate of instance b is the `m` set (om, cm, dm) when phase[b] = GF_PHASE_ACCEPT_MID, the `f` set when
 * GF_PHASE_ACCEPT_FINAL.  filt[B,cap,2] / nfilt[B] hold each instance's filter (PenaltyFilter.entries, penalty.py:176);
 * a dominated candidate sets phase[b] = GF_PHASE_REJECT and rho_pen[b] *= 10 (penalty.py:209-210), otherwise the entry is
 * inserted (dominated entries dropped) and rho[b] = rho_pen[b] (solver.py:364-369); overflow[b] = 1 if cap is too small. */
int gf_pareto_update(int B, int n, int m, const double* grad, const double* cons, const double* jty, const double* jtc,
                     const int32_t* phase, const int32_t* status, double opt_tol, double local_infeas_tol, double* rho,
                     const int32_t* work, const int32_t* nwork_dev, int nwork, void* stream);
int gf_filter_update(int B, int n, int m, int kind, int32_t* phase, const int32_t* status, const double* om,
                     const double* cm, const double* dm, const double* of, const double* cf, const double* df,
                     double* rho, double* rho_pen, double* filt, int32_t* nfilt, int cap, int32_t* overflow,
                     const int32_t* work, const int32_t* nwork_dev, int nwork, void* stream);

/* ---- fused persistent solver for small bound-constrained instances (cfg2) ----
 * Solver.solve (solver.py:233-431, default Params: DistanceRatio controller, simplified Newton, Symmetric step solver)
 * of the chained-Rosenbrock family with n <= 64 variables, bounds only: ONE WARP per instance runs the whole outer loop
 * -- termination test, residual + active set, Hessian, reduced system (tridiagonal for this family: LAPACK-style
 * dgttrf / dgttrs per run of inactive variables), both Newton steps, controller, commit -- for up to max_outer outer
 * iterations per launch; the state (x, grad, obj, lamb, err_sum, status, iters, accepted, newton_steps) lives in the
 * caller's arrays, so the call is repeated until every status is non-zero.  fresh != 0: evaluate grad / obj at x first.
 * params_host (HOST pointer, 12 doubles): opt_tol, active_tol, obj_lower_limit, newton_tol, lamb_red, lamb_min, lamb_max,
 * lamb_inc, theta_max, log(theta_ref), K_P, K_I (params.py:197-265); iteration_limit < 0 = none.
 * active[B,n] (may be NULL) receives the active set of each instance's last factorisation. */
int gf_rosen_fused_solve(int B, int n, const double* a, const double* b, const double* lb, const double* ub, double* x,
                         double* grad, double* obj, double* lamb, double* err_sum, int32_t* status, int32_t* iters,
                         int32_t* accepted, int32_t* newton_steps, double* total_res, uint8_t* active,
                         const double* params_host, int iteration_limit, int max_outer, int fresh, const int32_t* work,
                         const int32_t* nwork_dev, int nwork, void* stream);

/* helpers of the batched driver: ordered compaction of { b in parent (or 0..B-1) : (lo <= key[b] <= hi) != invert } */
int gf_build_worklist(int B, const int32_t* key, int lo, int hi, int invert, const int32_t* parent,
                      const int32_t* parent_count, int32_t* list, int32_t* count, void* stream);
int gf_dt_from_lamb(int B, const double* lamb, double* dt, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GRADFLOW_B200_H */
