// Residual / active set (a4, a5), rhs split and dual back-transform (a7), KKT assembly with
// active-set gather (a9-a11), step finish (a15), step-size control (a18), termination + penalty (a19).
//
// Reference (paths under /root/reference/pygradflow/):
//   implicit_func.py:21-60,150-161,219-252   residual F, projection, active set
//   step/solver/scaled_step_solver.py:38-107 b0/b1/b2 split, fact, dy back-transform
//   step/solver/symmetric_step_solver.py:27-94 H+lamb I, reduced symmetric K, reduced rhs
//   step/solver/step_solver.py:16-63         x+ = clip(x - dx), dx fix-up, diff
//   step/distance_ratio_control.py:18-78, controller.py:29-77, step/step_control.py:64-107
//   solver.py:180-205,323-326,357-378, iterate.py:115-181, active_set.py:4-29, penalty.py:46-74
#include "gf_common.cuh"
#include "../../include/gradflow_b200.h"

namespace {

// ---------------------------------------------------------------------------------------------
// F(x, y; x^, y^) and the active set.  scaled != 0: ScaledImplicitFunc, else ImplicitFunc.
// active_mode 0: recompute A from p (and store it when `active` != NULL); 1: use the stored A.
__global__ void residual_kernel(int n, int m, const double* __restrict__ x, const double* __restrict__ y,
                                const double* __restrict__ x0, const double* __restrict__ y0,
                                const double* __restrict__ dL, const double* __restrict__ cons,
                                const double* __restrict__ lb, const double* __restrict__ ub,
                                const double* __restrict__ dt, const double* __restrict__ tau, int scaled,
                                int active_mode, uint8_t* __restrict__ active, double* __restrict__ F,
                                double* __restrict__ nrm, GfWork work) {
    const int b = gf_instance(work, blockIdx.x);
    if (b < 0) return;
    __shared__ double red[32];
    const double dtb = dt[b];
    const double lamb = 1.0 / dtb;  // implicit_func.py:212
    // tau-variant of the point whose box position decides the active set (implicit_func.py:237-244,
    // newton_control.py:60-88); tau never enters the residual itself (:219-231)
    const bool use_tau = tau != nullptr && active_mode == 0;
    const double tb = use_tau ? tau[b] : 0.0;
    const double f_x = __dmul_rn(lamb, __dsub_rn(1.0, __dmul_rn(tb, lamb)));
    const double f_x0 = __dmul_rn(__dmul_rn(tb, lamb), lamb);
    const double f_d = __dmul_rn(tb, lamb);
    double ss = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const size_t o = (size_t)b * n + i;
        double p, lo, hi, xs;
        if (scaled) {
            p = __dsub_rn(__dmul_rn(lamb, x0[o]), dL[o]);  // :246
            lo = __dmul_rn(lamb, lb[o]);                    // :215-216
            hi = __dmul_rn(lamb, ub[o]);
            xs = __dmul_rn(lamb, x[o]);
        } else {
            p = __dsub_rn(x0[o], __dmul_rn(dtb, dL[o]));  // :148
            lo = lb[o];
            hi = ub[o];
            xs = x[o];
        }
        bool act;
        if (active_mode == 0) {
            double pa = p;
            if (use_tau) {
                if (scaled)  // f_x x + f_x0 x0 - f_d dL
                    pa = __dsub_rn(__dadd_rn(__dmul_rn(f_x, x[o]), __dmul_rn(f_x0, x0[o])), __dmul_rn(f_d, dL[o]));
                else         // (1 - tau lamb) x + (tau lamb) x0 - tau dL   (:133-146)
                    pa = __dsub_rn(__dadd_rn(__dmul_rn(__dsub_rn(1.0, f_d), x[o]), __dmul_rn(f_d, x0[o])),
                                   __dmul_rn(tb, dL[o]));
            }
            act = (pa < lo - GF_ACTIVE_SLACK) || (pa > hi + GF_ACTIVE_SLACK);  // :44
            if (active != nullptr) active[o] = act ? 1 : 0;
        } else {
            act = active[o] != 0;
        }
        const double proj = act ? fmin(fmax(p, lo), hi) : p;  // :46-60
        const double rx = __dsub_rn(xs, proj);
        if (F != nullptr) F[(size_t)b * (n + m) + i] = rx;
        ss += rx * rx;
    }
    for (int j = threadIdx.x; j < m; j += blockDim.x) {
        const size_t o = (size_t)b * m + j;
        double ry;
        if (scaled) {
            // -(lamb*y - (lamb*y0 + c))   :229
            ry = -__dsub_rn(__dmul_rn(lamb, y[o]), __dadd_rn(__dmul_rn(lamb, y0[o]), cons[o]));
        } else {
            ry = __dsub_rn(y[o], __dadd_rn(y0[o], __dmul_rn(dtb, cons[o])));  // :160
        }
        if (F != nullptr) F[(size_t)b * (n + m) + n + j] = ry;
        ss += ry * ry;
    }
    if (nrm != nullptr) {
        const double s = block_sum(ss, red);
        if (threadIdx.x == 0) nrm[b] = sqrt(s);
    }
}

// ---------------------------------------------------------------------------------------------
// perm[b] = (inactive indices ascending, then active indices ascending); nI[b] = #inactive.
__global__ void index_sets_kernel(int n, int m, const uint8_t* __restrict__ active, int32_t* __restrict__ perm,
                                  int32_t* __restrict__ nI, int32_t* __restrict__ Nvec, GfWork work) {
    const int b = gf_instance(work, blockIdx.x);
    if (b < 0) return;
    __shared__ int cnt[1025];
    const int T = blockDim.x;
    const int per = (n + T - 1) / T;
    const int lo = threadIdx.x * per, hi = min(n, lo + per);
    const uint8_t* ab = active + (size_t)b * n;
    int c = 0;
    for (int i = lo; i < hi; i++) c += ab[i] ? 0 : 1;
    cnt[threadIdx.x + 1] = c;
    if (threadIdx.x == 0) cnt[0] = 0;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int t = 1; t <= T; t++) cnt[t] += cnt[t - 1];
    }
    __syncthreads();
    const int total_inactive = cnt[T];
    int pi = cnt[threadIdx.x];                      // inactive before my segment
    int pa = total_inactive + (min(lo, n) - pi);    // actives go after all inactives
    int32_t* pb = perm + (size_t)b * n;
    for (int i = lo; i < hi; i++) {
        if (ab[i]) pb[pa++] = i; else pb[pi++] = i;
    }
    if (threadIdx.x == 0) {
        nI[b] = total_inactive;
        if (Nvec != nullptr) Nvec[b] = total_inactive + m;
    }
}

// ---------------------------------------------------------------------------------------------
// K = [[H_II + lamb I, J_I'], [J_I, -lamb/(1+lamb rho) I]] of order N = nI + m in the top-left corner of
// K[b] (ld x ld, row-major); rows/cols N..Npad-1 (Npad = N rounded up to `pad`) are identity so blocked
// factorisations need no edge handling.  lower_only: write only c <= r (LDL' path).
template <int ROWS>
__global__ void kkt_assemble_kernel(int n, int m, int ld, int pad, int lower_only, const double* __restrict__ H,
                                    const double* __restrict__ J, const int32_t* __restrict__ perm,
                                    const int32_t* __restrict__ nIv, const double* __restrict__ dt,
                                    const double* __restrict__ rho, double* __restrict__ K, GfWork work) {
    const int b = gf_instance(work, blockIdx.y);
    if (b < 0) return;
    const int nI = nIv[b];
    const int N = nI + m;
    const int Npad = min(ld, ((N + pad - 1) / pad) * pad);
    const int r0 = blockIdx.x * ROWS;
    if (r0 >= Npad) return;
    extern __shared__ int32_t sperm[];
    for (int i = threadIdx.x; i < nI; i += blockDim.x) sperm[i] = perm[(size_t)b * n + i];
    __syncthreads();
    const double lamb = 1.0 / dt[b];                       // symmetric_step_solver.py:30
    const double corner = -lamb / (1.0 + lamb * rho[b]);   // :60-62
    const double* Hb = H + (size_t)b * n * n;
    const double* Jb = J + (size_t)b * m * n;
    double* Kb = K + (size_t)b * ld * ld;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int rr = wid; rr < ROWS; rr += nw) {
        const int r = r0 + rr;
        if (r >= Npad) break;
        double* out = Kb + (size_t)r * ld;
        const int cend = lower_only ? (r + 1) : Npad;
        if (r < nI) {
            const double* src = Hb + (size_t)sperm[r] * n;
            const int c1 = min(cend, nI);
            for (int c = lane; c < c1; c += 32) {
                double v = __ldg(src + sperm[c]);
                if (c == r) v = __dadd_rn(v, lamb);  // H + diag(lamb)  :34-36
                out[c] = v;
            }
            if (!lower_only) {
                const int pr = sperm[r];
                for (int c = nI + lane; c < N; c += 32) out[c] = __ldg(Jb + (size_t)(c - nI) * n + pr);
                for (int c = N + lane; c < Npad; c += 32) out[c] = 0.0;
            }
        } else if (r < N) {
            const double* src = Jb + (size_t)(r - nI) * n;
            for (int c = lane; c < nI; c += 32) out[c] = __ldg(src + sperm[c]);
            for (int c = nI + lane; c < cend; c += 32) out[c] = (c == r) ? corner : 0.0;
        } else {
            for (int c = lane; c < cend; c += 32) out[c] = (c == r) ? 1.0 : 0.0;
        }
    }
}

// rhs = [ b1 - H_lamb[I,A] b0 ; b2t - J[:,A] b0 ],  b0 = dt rx[A], b1 = rx[I], b2t = fact ry.
// H[I,A] is read as H[A,I] (H symmetric) so each active row streams coalesced; the sum runs over the
// active indices in ascending order with un-fused mul/add (scipy csc_matvec order).
__global__ void kkt_rhs_kernel(int n, int m, int ld, const double* __restrict__ H, const double* __restrict__ J,
                               const int32_t* __restrict__ perm, const int32_t* __restrict__ nIv,
                               const double* __restrict__ F, const double* __restrict__ dt,
                               const double* __restrict__ rho, double* __restrict__ rhs, GfWork work) {
    const int b = gf_instance(work, blockIdx.x);
    if (b < 0) return;
    extern __shared__ unsigned char smraw[];
    const int nI = nIv[b];
    const int nA = n - nI;
    double* b0 = reinterpret_cast<double*>(smraw);         // n doubles (nA used)
    int32_t* sp = reinterpret_cast<int32_t*>(b0 + n);      // n ints
    const double dtb = dt[b];
    const double lamb = 1.0 / dtb;
    const double fact = 1.0 / (1.0 + lamb * rho[b]);       // scaled_step_solver.py:93
    const double* Fb = F + (size_t)b * (n + m);
    for (int i = threadIdx.x; i < n; i += blockDim.x) sp[i] = perm[(size_t)b * n + i];
    __syncthreads();
    for (int k = threadIdx.x; k < nA; k += blockDim.x) b0[k] = __dmul_rn(dtb, Fb[sp[nI + k]]);  // :56
    __syncthreads();
    const double* Hb = H + (size_t)b * n * n;
    double* out = rhs + (size_t)b * ld;
    for (int r = threadIdx.x; r < nI; r += blockDim.x) {
        const int pr = sp[r];
        double acc = 0.0;
        for (int k = 0; k < nA; k++) acc = __dadd_rn(acc, __dmul_rn(__ldg(Hb + (size_t)sp[nI + k] * n + pr), b0[k]));
        out[r] = __dsub_rn(Fb[pr], acc);  // symmetric_step_solver.py:89
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const double* Jb = J + (size_t)b * m * n;
    for (int j = wid; j < m; j += nw) {
        const double* row = Jb + (size_t)j * n;
        double acc = 0.0;
        for (int k = lane; k < nA; k += 32) acc += __ldg(row + sp[nI + k]) * b0[k];
        acc = warp_sum(acc);
        if (lane == 0) out[nI + j] = __dsub_rn(__dmul_rn(fact, Fb[n + j]), acc);  // :90
    }
    for (int c = nI + m + threadIdx.x; c < ld; c += blockDim.x) out[c] = 0.0;
}

// ---------------------------------------------------------------------------------------------
// The unsymmetric full-order formulations of the scaled step (step/solver/__init__.py:12-31), order n + m, columns in
// the natural order (x, y):
//   GF_FORM_ASYMMETRIC (asymmetric_step_solver.py:77-104): rows of [[H + lamb I, J'], [J, -lamb fact I]] in the
//       natural order, the row of every active variable overwritten by the unit row (:37-75);
//   GF_FORM_EXTENDED (extended_step_solver.py:39-83): (selector rows of the active variables, ascending; the
//       inactive rows of [H + lamb I, J'], ascending; [J, -lamb fact I]);
//   GF_FORM_STANDARD (standard_step_solver.py:40-53 + implicit_func.py:163-199): the derivative of the UNSCALED implicit
//       function, [[I + P_I dt H, P_I dt J'], [-dt J, I]] with H = H_rho handed in by the caller.
// One warp per row, rows of K written coalesced.
template <int ROWS>
__global__ void kkt_full_kernel(int n, int m, int ld, int form, const double* __restrict__ H,
                                const double* __restrict__ J, const int32_t* __restrict__ perm,
                                const int32_t* __restrict__ nIv, const uint8_t* __restrict__ active,
                                const double* __restrict__ dt, const double* __restrict__ rho, double* __restrict__ K,
                                GfWork work) {
    const int b = gf_instance(work, blockIdx.y);
    if (b < 0) return;
    const int N = n + m;
    const int r0 = blockIdx.x * ROWS;
    if (r0 >= N) return;
    const int nI = nIv[b], nA = n - nI;
    const double lamb = 1.0 / dt[b];
    const double corner = -lamb / __dadd_rn(1.0, __dmul_rn(lamb, rho[b]));
    const double* Hb = H + (size_t)b * n * n;
    const double* Jb = J + (size_t)b * m * n;
    const int32_t* pb = perm + (size_t)b * n;
    const uint8_t* ab = active + (size_t)b * n;
    double* Kb = K + (size_t)b * ld * ld;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int rr = wid; rr < ROWS; rr += nw) {
        const int r = r0 + rr;
        if (r >= N) break;
        double* out = Kb + (size_t)r * ld;
        const double dtb = dt[b];
        if (r >= n) {
            const double* src = Jb + (size_t)(r - n) * n;
            if (form == GF_FORM_STANDARD) {  // [-dt J, I]  (implicit_func.py:185-186)
                for (int c = lane; c < n; c += 32) out[c] = -__dmul_rn(dtb, __ldg(src + c));
                for (int c = n + lane; c < N; c += 32) out[c] = (c == r) ? 1.0 : 0.0;
            } else if (form == GF_FORM_SCALED_DERIV) {  // [-J, lamb I]  (implicit_func.py:274-276)
                for (int c = lane; c < n; c += 32) out[c] = -__ldg(src + c);
                for (int c = n + lane; c < N; c += 32) out[c] = (c == r) ? lamb : 0.0;
            } else {
                for (int c = lane; c < n; c += 32) out[c] = __ldg(src + c);
                for (int c = n + lane; c < N; c += 32) out[c] = (c == r) ? corner : 0.0;
            }
            continue;
        }
        int unit = -1, hrow = -1;  // unit row e_unit, or row hrow of [H + lamb I, J']
        if (form == GF_FORM_EXTENDED) {
            if (r < nA) unit = pb[nI + r];
            else hrow = pb[r - nA];
        } else {
            if (ab[r]) unit = r;
            else hrow = r;
        }
        if (unit >= 0) {  // SCALED_DERIV: keep_rows zeroes the row of H_rho and J', lamb I stays (implicit_func.py:266-270)
            const double one = form == GF_FORM_SCALED_DERIV ? lamb : 1.0;
            for (int c = lane; c < N; c += 32) out[c] = (c == unit) ? one : 0.0;
        } else if (form == GF_FORM_STANDARD) {  // [I + dt H_rho, dt J'] on the inactive rows (implicit_func.py:176-183)
            const double* src = Hb + (size_t)hrow * n;
            for (int c = lane; c < n; c += 32) {
                double v = __dmul_rn(dtb, __ldg(src + c));
                if (c == hrow) v = __dadd_rn(1.0, v);
                out[c] = v;
            }
            for (int c = lane; c < m; c += 32) out[n + c] = __dmul_rn(dtb, __ldg(Jb + (size_t)c * n + hrow));
        } else {
            const double* src = Hb + (size_t)hrow * n;
            for (int c = lane; c < n; c += 32) {
                double v = __ldg(src + c);
                if (c == hrow) v = __dadd_rn(v, lamb);
                out[c] = v;
            }
            for (int c = lane; c < m; c += 32) out[n + c] = __ldg(Jb + (size_t)c * n + hrow);
        }
    }
}

// rhs of the full-order formulations: b0 = dt F_x[A], b1 = F_x[I], b2t = fact F_y (scaled_step_solver.py:38-60,91-97)
// placed by position (asymmetric_step_solver.py:106-123) or as (b0, b1, b2t) (extended_step_solver.py:93).
__global__ void kkt_full_rhs_kernel(int n, int m, int ld, int form, const int32_t* __restrict__ perm,
                                    const int32_t* __restrict__ nIv, const uint8_t* __restrict__ active,
                                    const double* __restrict__ F, const double* __restrict__ dt,
                                    const double* __restrict__ rho, double* __restrict__ rhs, GfWork work) {
    const int b = gf_instance(work, blockIdx.x);
    if (b < 0) return;
    const int nI = nIv[b], nA = n - nI;
    const double dtb = dt[b];
    const double lamb = 1.0 / dtb;
    const double fact = 1.0 / __dadd_rn(1.0, __dmul_rn(lamb, rho[b]));
    const double* Fb = F + (size_t)b * (n + m);
    const int32_t* pb = perm + (size_t)b * n;
    const uint8_t* ab = active + (size_t)b * n;
    double* out = rhs + (size_t)b * ld;
    if (form == GF_FORM_STANDARD) {  // rhs = F of the unscaled implicit function (standard_step_solver.py:63)
        for (int r = threadIdx.x; r < n + m; r += blockDim.x) out[r] = Fb[r];
        for (int c = n + m + threadIdx.x; c < ld; c += blockDim.x) out[c] = 0.0;
        return;
    }
    for (int r = threadIdx.x; r < n; r += blockDim.x) {
        if (form == GF_FORM_EXTENDED) out[r] = r < nA ? __dmul_rn(dtb, Fb[pb[nI + r]]) : Fb[pb[r - nA]];
        else out[r] = ab[r] ? __dmul_rn(dtb, Fb[r]) : Fb[r];
    }
    for (int j = threadIdx.x; j < m; j += blockDim.x) out[n + j] = __dmul_rn(fact, Fb[n + j]);
    for (int c = n + m + threadIdx.x; c < ld; c += blockDim.x) out[c] = 0.0;
}

// dx[I] = s[:nI], dx[A] = b0, dy = fact (sy - rho b2); x+ = clip(x - dx) with dx fix-up; y+ = y - dy;
// diff = sqrt(|dx|^2 + |dy|^2) with the post-clip dx.
__global__ void step_finish_kernel(int n, int m, int ld, const double* __restrict__ xbase,
                                   const double* __restrict__ ybase, const double* __restrict__ sol,
                                   const int32_t* __restrict__ perm, const int32_t* __restrict__ nIv,
                                   const double* __restrict__ F, const double* __restrict__ dt,
                                   const double* __restrict__ rho, const double* __restrict__ lb,
                                   const double* __restrict__ ub, double* __restrict__ xn, double* __restrict__ yn,
                                   double* __restrict__ dxo, double* __restrict__ dyo, double* __restrict__ diff,
                                   GfWork work) {
    const int b = gf_instance(work, blockIdx.x);
    if (b < 0) return;
    __shared__ double red[32];
    const int nI = nIv[b];
    const double dtb = dt[b];
    const double lamb = 1.0 / dtb;
    const double rh = rho[b];
    const double fact = 1.0 / (1.0 + lamb * rh);
    const double* Fb = F + (size_t)b * (n + m);
    const double* s = sol + (size_t)b * ld;
    double ss = 0.0;
    for (int r = threadIdx.x; r < n; r += blockDim.x) {
        const int j = perm[(size_t)b * n + r];
        const size_t o = (size_t)b * n + j;
        double dx = (r < nI) ? s[r] : __dmul_rn(dtb, Fb[j]);
        const double xj = xbase[o];
        double v = __dsub_rn(xj, dx);
        const double l = lb[o], u = ub[o];
        if (v < l) { v = l; dx = __dsub_rn(xj, l); }   // step_solver.py:35-37
        if (v > u) { v = u; dx = __dsub_rn(xj, u); }   // :39-41
        xn[o] = v;
        if (dxo != nullptr) dxo[o] = dx;
        ss += dx * dx;
    }
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
        const size_t o = (size_t)b * m + i;
        const double dy = __dmul_rn(fact, __dsub_rn(s[nI + i], __dmul_rn(rh, Fb[n + i])));  // scaled_step_solver.py:105
        yn[o] = __dsub_rn(ybase[o], dy);
        if (dyo != nullptr) dyo[o] = dy;
        ss += dy * dy;
    }
    const double t = block_sum(ss, red);
    if (threadIdx.x == 0) diff[b] = sqrt(t);
}

// ---------------------------------------------------------------------------------------------
// Solver._check_terminate for every running instance (status == 0).
__global__ void terminate_kernel(int n, int m, const double* __restrict__ x, const double* __restrict__ grad,
                                 const double* __restrict__ cons, const double* __restrict__ jty,
                                 const double* __restrict__ jtc, const double* __restrict__ obj,
                                 const double* __restrict__ lb, const double* __restrict__ ub, double opt_tol,
                                 double active_tol, double local_infeas_tol, double obj_lower_limit,
                                 int iteration_limit, const int32_t* __restrict__ iters,
                                 int32_t* __restrict__ status, double* __restrict__ total_res, GfWork work) {
    const int b = gf_instance(work, blockIdx.x);
    if (b < 0) return;
    if (status[b] != 0) return;
    __shared__ double red[32];
    double cv = 0.0, bv = 0.0, st = 0.0, li = 0.0;
    for (int j = threadIdx.x; j < m; j += blockDim.x) cv = fmax(cv, fabs(cons[(size_t)b * m + j]));
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const size_t o = (size_t)b * n + i;
        const double xi = x[o], l = lb[o], u = ub[o];
        bv = fmax(bv, fmax(fmax(l - xi, 0.0), fmax(xi - u, 0.0)));       // iterate.py:151-160
        const bool atl = fabs(xi - l) <= active_tol;                      // active_set.py:11-12
        const bool atu = fabs(u - xi) <= active_tol;
        const bool both = atl && atu, lo = atl && !both, up = atu && !both;
        const double gj = __dadd_rn(grad[o], m > 0 ? jty[o] : 0.0);
        const double r = -gj;
        double d = 0.0;
        if (up) d = fmax(r, 0.0);                                         // iterate.py:143-145
        if (lo) d = fmin(r, 0.0);
        if (both) d = r;
        st = fmax(st, fabs(__dadd_rn(gj, d)));                            // :171-172
        if (m > 0) {
            double q = jtc[o];
            if (lo) q = fmin(q, 0.0);                                     // :130-131
            if (up) q = fmax(q, 0.0);
            li = fmax(li, fabs(q));
        }
    }
    cv = block_max(cv, red);
    bv = block_max(bv, red);
    st = block_max(st, red);
    li = block_max(li, red);
    if (threadIdx.x == 0) {
        const double tot = fmax(cv, fmax(bv, st));
        if (total_res != nullptr) total_res[b] = tot;
        int s = 0;
        if (iteration_limit >= 0 && iters[b] >= iteration_limit) s = GF_STATUS_ITERATION_LIMIT;  // solver.py:183-187
        else if (tot <= opt_tol) s = GF_STATUS_OPTIMAL;                                           // :193-195
        else if (cv > opt_tol && li <= local_infeas_tol) s = GF_STATUS_LOCALLY_INFEASIBLE;        // :197-199
        else if (obj[b] <= obj_lower_limit && cv <= opt_tol && bv <= opt_tol) s = GF_STATUS_UNBOUNDED;  // :201-205
        status[b] = s;
    }
}

// DistanceRatioController.step, decision after the first Newton step (distance_ratio_control.py:27-44)
// and the solver-failure path of StepController.compute_step (step_control.py:80-83,102-104).
__global__ void dr_first_kernel(int B, const int32_t* __restrict__ status, const int32_t* __restrict__ info,
                                const double* __restrict__ dt, const double* __restrict__ mid_norm,
                                const double* __restrict__ diff1, double newton_tol, double lamb_red,
                                double lamb_min, int32_t* __restrict__ phase, double* __restrict__ lamb_next) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    if (status[b] != 0) { phase[b] = GF_PHASE_IDLE; return; }
    const double lamb = 1.0 / dt[b];
    if (info != nullptr && info[b] != 0) {
        phase[b] = GF_PHASE_FAILED;
        lamb_next[b] = 2.0 * lamb;
    } else if (mid_norm[b] <= newton_tol) {
        phase[b] = GF_PHASE_ACCEPT_MID;
        lamb_next[b] = fmax(lamb * lamb_red, lamb_min);
    } else if (diff1[b] == 0.0) {
        phase[b] = GF_PHASE_ACCEPT_MID;
        lamb_next[b] = lamb;
    } else {
        phase[b] = GF_PHASE_SECOND;
    }
}

// Decision after the second Newton step (distance_ratio_control.py:46-78) with the log-PI controller
// (controller.py:44-77; its integral term is never reset -- SURVEY 7).
__global__ void dr_second_kernel(int B, const double* __restrict__ dt, const double* __restrict__ diff1,
                                 const double* __restrict__ diff2, double theta_max, double log_theta_ref,
                                 double K_P, double K_I, double lamb_min, double lamb_inc,
                                 double* __restrict__ err_sum, int32_t* __restrict__ phase,
                                 double* __restrict__ lamb_next, double* __restrict__ theta_out) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    if (phase[b] != GF_PHASE_SECOND) return;
    const double lamb = 1.0 / dt[b];
    if (diff2[b] == 0.0) {
        phase[b] = GF_PHASE_ACCEPT_FINAL;
        lamb_next[b] = lamb;
        return;
    }
    const double theta = diff2[b] / diff1[b];
    if (theta_out != nullptr) theta_out[b] = theta;
    if (theta <= theta_max) {
        const double err = log_theta_ref - log(theta);
        const double es = err_sum[b] + err;
        err_sum[b] = es;
        const double mod = exp(K_P * err + K_I * es);
        lamb_next[b] = fmax(lamb_min, lamb / mod);
        phase[b] = GF_PHASE_ACCEPT_FINAL;
    } else {
        lamb_next[b] = lamb * lamb_inc;
        phase[b] = GF_PHASE_REJECT;
    }
}


// ResiduumRatioController.step (residuum_ratio_control.py:18-63) / FixedStepSizeController.step (fixed_control.py:12-19)
// after their single Newton step, with the solver-failure path of StepController.compute_step (step_control.py:102-104).
__global__ void single_control_kernel(int B, int fixed, const int32_t* __restrict__ status,
                                      const int32_t* __restrict__ info, const double* __restrict__ dt,
                                      const double* __restrict__ mid_norm, const double* __restrict__ orig_norm,
                                      double newton_tol, double theta_max, double log_theta_ref, double K_P, double K_I,
                                      double lamb_red, double lamb_min, double lamb_inc, double lamb_init,
                                      double* __restrict__ err_sum, int32_t* __restrict__ phase,
                                      double* __restrict__ lamb_next, double* __restrict__ theta_out,
                                      unsigned long long* __restrict__ nsteps) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    if (status[b] != 0) { phase[b] = GF_PHASE_IDLE; return; }
    const double lamb = 1.0 / dt[b];
    if (info[b] != 0) {
        phase[b] = GF_PHASE_FAILED;
        lamb_next[b] = 2.0 * lamb;
        return;
    }
    atomicAdd(nsteps, 1ULL);
    if (fixed) {
        phase[b] = GF_PHASE_ACCEPT_MID;
        lamb_next[b] = lamb_init;
        return;
    }
    if (mid_norm[b] <= newton_tol) {                       // residuum_ratio_control.py:33-38
        phase[b] = GF_PHASE_ACCEPT_MID;
        lamb_next[b] = fmax(lamb * lamb_red, lamb_min);
        return;
    }
    const double theta = mid_norm[b] / orig_norm[b];
    theta_out[b] = theta;
    if (theta <= theta_max) {                              // :50-58, controller.py:44-77
        const double err = log_theta_ref - log(theta);
        const double es = err_sum[b] + err;
        err_sum[b] = es;
        const double mod = exp(K_P * err + K_I * es);
        lamb_next[b] = fmax(lamb / mod, lamb_min);
        phase[b] = GF_PHASE_ACCEPT_MID;
    } else {
        lamb_next[b] = lamb * lamb_inc;
        phase[b] = GF_PHASE_REJECT;
    }
}

// ExactController.step (exact_control.py:16-66), one call per stage of its Newton loop.  `live` = instances still
// inside the loop.  mode 0: after the first step (failure => FAILED, 2 lambda; curr = |F(orig)|); mode 1: look at the
// residual norm `val` of Newton iterate number `it` (0-based; even iterates live in `mid` -> phase 2, odd in `fin` ->
// phase 3): converged => accept, lambda / 2; contraction rate > rate_bound or the last iterate => reject, 2 lambda;
// mode 2: a failed refactorisation after a further step ends the loop like a StepSolverError.
__global__ void exact_control_kernel(int B, int mode, int it, int last, const int32_t* __restrict__ status,
                                     const int32_t* __restrict__ info, const double* __restrict__ dt,
                                     const double* __restrict__ val, const double* __restrict__ orig_norm,
                                     double newton_tol, double rate_bound, double* __restrict__ curr,
                                     int32_t* __restrict__ live, int32_t* __restrict__ phase,
                                     double* __restrict__ lamb_next, unsigned long long* __restrict__ nsteps) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const double lamb = 1.0 / dt[b];
    if (mode == 0) {
        const bool running = status[b] == 0;
        const bool failed = running && info[b] != 0;
        live[b] = (running && !failed) ? 1 : 0;
        curr[b] = orig_norm[b];
        phase[b] = failed ? GF_PHASE_FAILED : GF_PHASE_IDLE;
        lamb_next[b] = failed ? 2.0 * lamb : lamb;
        return;
    }
    if (!live[b]) return;
    if (mode == 2) {
        if (info[b] != 0) {
            phase[b] = GF_PHASE_FAILED;
            lamb_next[b] = 2.0 * lamb;
            live[b] = 0;
        }
        return;
    }
    atomicAdd(nsteps, 1ULL);
    const double v = val[b];
    if (v <= newton_tol) {                                  // exact_control.py:44-49
        phase[b] = GF_PHASE_ACCEPT_MID + (it & 1);
        lamb_next[b] = 0.5 * lamb;
        live[b] = 0;
    } else if ((v / curr[b]) > rate_bound || last) {        // :51-58, :60-66
        phase[b] = GF_PHASE_REJECT;
        lamb_next[b] = 2.0 * lamb;
        live[b] = 0;
    } else {
        curr[b] = v;
    }
}

// End of an outer iteration (solver.py:318-378): lambda hand-over, lamb_max guard, penalty update
// (penalty.py:36-113; dual_norm_update = 0 constant, 1 DualNorm, 2 DualEquilibration), iterate <- accepted Newton
// iterate, counters.
__global__ void commit_kernel(int n, int m, const int32_t* __restrict__ phase, const double* __restrict__ lamb_next,
                              double lamb_max, int dual_norm_update, const double* __restrict__ xm,
                              const double* __restrict__ ym, const double* __restrict__ gm,
                              const double* __restrict__ cm, const double* __restrict__ om,
                              const double* __restrict__ xf, const double* __restrict__ yf,
                              const double* __restrict__ gf, const double* __restrict__ cf,
                              const double* __restrict__ of, double* __restrict__ x, double* __restrict__ y,
                              double* __restrict__ grad, double* __restrict__ cons, double* __restrict__ obj,
                              double* __restrict__ lamb, double* __restrict__ rho, int32_t* __restrict__ iters,
                              int32_t* __restrict__ accepted, int32_t* __restrict__ status) {
    const int b = blockIdx.x;
    const int ph = phase[b];
    if (ph == GF_PHASE_IDLE || status[b] != 0) return;
    __shared__ double red[32];
    const double ln = lamb_next[b];
    if (ln >= lamb_max) {  // solver.py:323-326 raises; the batched driver records a status instead
        if (threadIdx.x == 0) { status[b] = GF_STATUS_LAMB_MAX; lamb[b] = ln; }
        return;
    }
    const bool acc_mid = ph == GF_PHASE_ACCEPT_MID, acc_fin = ph == GF_PHASE_ACCEPT_FINAL;
    if (acc_mid || acc_fin) {
        const double* xs = acc_mid ? xm : xf;
        const double* ys = acc_mid ? ym : yf;
        const double* gs = acc_mid ? gm : gf;
        const double* cs = acc_mid ? cm : cf;
        double ymax = 0.0, yc = 0.0, cc = 0.0;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const size_t o = (size_t)b * n + i;
            x[o] = xs[o];
            grad[o] = gs[o];
        }
        for (int j = threadIdx.x; j < m; j += blockDim.x) {
            const size_t o = (size_t)b * m + j;
            const double yy = ys[o], cj = cs[o];
            y[o] = yy;
            cons[o] = cj;
            ymax = fmax(ymax, fabs(yy));
            yc += yy * cj;
            cc += cj * cj;
        }
        ymax = block_max(ymax, red);
        if (dual_norm_update == 2) {  // uniform per launch
            yc = block_sum(yc, red);
            cc = block_sum(cc, red);
        }
        if (threadIdx.x == 0) {
            obj[b] = acc_mid ? om[b] : of[b];
            const double r = rho[b];
            if (dual_norm_update == 1 && m > 0) {          // DualNormUpdate, penalty.py:59-74
                if (ymax >= 10.0 * r) rho[b] = fmin(ymax, 10.0 * r);
            } else if (dual_norm_update == 2) {            // DualEquilibration, penalty.py:90-113
                const double viol = 0.5 * cc;
                if (viol != 0.0) {
                    const double target = 0.01 * fabs(yc) / viol;
                    if (r < target) rho[b] = fmax(r * 10.0, target);
                }
            }
            accepted[b] += 1;
        }
    }
    if (threadIdx.x == 0) {
        lamb[b] = ln;
        iters[b] += 1;
    }
}

// Ordered compaction over a parent list (or 0..B-1): out = { b in parent : (lo <= key[b] <= hi) != invert },
// order preserved, *count = |out|.  Single CTA.
__global__ void worklist_kernel(int B, const int32_t* __restrict__ key, int lo, int hi, int invert,
                                const int32_t* __restrict__ parent, const int32_t* __restrict__ parent_count,
                                int32_t* __restrict__ list, int32_t* __restrict__ count) {
    __shared__ int cnt[1025];
    const int T = blockDim.x;
    const int total = parent_count != nullptr ? *parent_count : B;
    const int per = (total + T - 1) / T;
    const int s = min(total, (int)threadIdx.x * per), e = min(total, s + per);
    int c = 0;
    for (int w = s; w < e; w++) {
        const int b = parent != nullptr ? parent[w] : w;
        const bool in = key[b] >= lo && key[b] <= hi;
        c += (in != (invert != 0)) ? 1 : 0;
    }
    cnt[threadIdx.x + 1] = c;
    if (threadIdx.x == 0) cnt[0] = 0;
    __syncthreads();
    if (threadIdx.x == 0) for (int t = 1; t <= T; t++) cnt[t] += cnt[t - 1];
    __syncthreads();
    int p = cnt[threadIdx.x];
    for (int w = s; w < e; w++) {
        const int b = parent != nullptr ? parent[w] : w;
        const bool in = key[b] >= lo && key[b] <= hi;
        if (in != (invert != 0)) list[p++] = b;
    }
    if (threadIdx.x == 0) *count = cnt[T];
}

__global__ void dt_from_lamb_kernel(int B, const double* __restrict__ lamb, double* __restrict__ dt) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) dt[b] = 1.0 / lamb[b];  // solver.py:315 passes 1.0 / lamb
}

inline int pick_threads(int n) {
    int t = ((n + 31) / 32) * 32;
    if (t < 64) t = 64;
    if (t > 512) t = 512;
    return t;
}

}  // namespace

extern "C" int gf_residual(int B, int n, int m, const double* x, const double* y, const double* x0, const double* y0,
                           const double* dL, const double* cons, const double* lb, const double* ub,
                           const double* dt, int scaled, int active_mode, uint8_t* active, double* F, double* nrm,
                           const int32_t* work, const int32_t* nwork_dev, int nwork, void* stream) {
    if (B <= 0 || n <= 0 || m < 0 || !x || !x0 || !dL || !lb || !ub || !dt) return GF_ERR_ARG;
    if (m > 0 && (!y || !y0 || !cons)) return GF_ERR_ARG;
    if (active_mode == 1 && !active) return GF_ERR_ARG;
    if (nwork <= 0) return GF_OK;
    residual_kernel<<<nwork, pick_threads(n + m), 0, (cudaStream_t)stream>>>(
        n, m, x, y, x0, y0, dL, cons, lb, ub, dt, nullptr, scaled, active_mode, active, F, nrm, GfWork{work, nwork_dev});
    return gf_launch_status();
}

extern "C" int gf_residual_tau(int B, int n, int m, const double* x, const double* y, const double* x0,
                               const double* y0, const double* dL, const double* cons, const double* lb,
                               const double* ub, const double* dt, const double* tau, int scaled, int active_mode,
                               uint8_t* active, double* F, double* nrm, const int32_t* work, const int32_t* nwork_dev,
                               int nwork, void* stream) {
    if (B <= 0 || n <= 0 || m < 0 || !x || !x0 || !dL || !lb || !ub || !dt) return GF_ERR_ARG;
    if (m > 0 && (!y || !y0 || !cons)) return GF_ERR_ARG;
    if (active_mode == 1 && !active) return GF_ERR_ARG;
    if (nwork <= 0) return GF_OK;
    residual_kernel<<<nwork, pick_threads(n + m), 0, (cudaStream_t)stream>>>(
        n, m, x, y, x0, y0, dL, cons, lb, ub, dt, tau, scaled, active_mode, active, F, nrm, GfWork{work, nwork_dev});
    return gf_launch_status();
}

extern "C" int gf_index_sets(int B, int n, int m, const uint8_t* active, int32_t* perm, int32_t* nI, int32_t* Nvec,
                             const int32_t* work, const int32_t* nwork_dev, int nwork, void* stream) {
    if (B <= 0 || n <= 0 || !active || !perm || !nI) return GF_ERR_ARG;
    if (nwork <= 0) return GF_OK;
    int threads = n >= 1024 ? 256 : (n >= 256 ? 128 : 64);
    index_sets_kernel<<<nwork, threads, 0, (cudaStream_t)stream>>>(n, m, active, perm, nI, Nvec,
                                                                    GfWork{work, nwork_dev});
    return gf_launch_status();
}

extern "C" int gf_kkt_assemble(int B, int n, int m, int ld, int pad, int lower_only, const double* H, const double* J,
                               const int32_t* perm, const int32_t* nI, const double* dt, const double* rho, double* K,
                               const int32_t* work, const int32_t* nwork_dev, int nwork, void* stream) {
    if (B <= 0 || n <= 0 || m < 0 || ld < n + m || pad < 1 || !H || !perm || !nI || !dt || !rho || !K)
        return GF_ERR_ARG;
    if (m > 0 && !J) return GF_ERR_ARG;
    if (nwork <= 0) return GF_OK;
    if (nwork > 65535) return GF_ERR_UNSUPPORTED;  // the instance index rides in grid.y
    constexpr int ROWS = 32;
    dim3 grid((ld + ROWS - 1) / ROWS, nwork);
    const size_t smem = (size_t)n * sizeof(int32_t);
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(kkt_assemble_kernel<ROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kkt_assemble_kernel<ROWS><<<grid, 256, smem, (cudaStream_t)stream>>>(n, m, ld, pad, lower_only, H, J, perm, nI, dt,
                                                                         rho, K, GfWork{work, nwork_dev});
    return gf_launch_status();
}

extern "C" int gf_kkt_rhs(int B, int n, int m, int ld, const double* H, const double* J, const int32_t* perm,
                          const int32_t* nI, const double* F, const double* dt, const double* rho, double* rhs,
                          const int32_t* work, const int32_t* nwork_dev, int nwork, void* stream) {
    if (B <= 0 || n <= 0 || m < 0 || ld < n + m || !H || !perm || !nI || !F || !dt || !rho || !rhs) return GF_ERR_ARG;
    if (m > 0 && !J) return GF_ERR_ARG;
    if (nwork <= 0) return GF_OK;
    const size_t smem = (size_t)n * (sizeof(double) + sizeof(int32_t));
    if (smem > 200 * 1024) return GF_ERR_UNSUPPORTED;
    if (smem > 48 * 1024) cudaFuncSetAttribute(kkt_rhs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kkt_rhs_kernel<<<nwork, pick_threads(n), smem, (cudaStream_t)stream>>>(n, m, ld, H, J, perm, nI, F, dt, rho, rhs,
                                                                            GfWork{work, nwork_dev});
    return gf_launch_status();
}

extern "C" int gf_kkt_assemble_full(int B, int n, int m, int ld, int form, const double* H, const double* J,
                                    const int32_t* perm, const int32_t* nI, const uint8_t* active, const double* dt,
                                    const double* rho, double* K, const int32_t* work, const int32_t* nwork_dev,
                                    int nwork, void* stream) {
    if (B <= 0 || n <= 0 || m < 0 || ld < n + m || !H || !perm || !nI || !active || !dt || !rho || !K) return GF_ERR_ARG;
    if (m > 0 && !J) return GF_ERR_ARG;
    if (form != GF_FORM_ASYMMETRIC && form != GF_FORM_EXTENDED && form != GF_FORM_STANDARD &&
        form != GF_FORM_SCALED_DERIV)
        return GF_ERR_UNSUPPORTED;
    if (nwork <= 0) return GF_OK;
    if (nwork > 65535) return GF_ERR_UNSUPPORTED;  // the instance index rides in grid.y
    constexpr int ROWS = 16;
    dim3 grid((n + m + ROWS - 1) / ROWS, nwork);
    kkt_full_kernel<ROWS><<<grid, 256, 0, (cudaStream_t)stream>>>(n, m, ld, form, H, J, perm, nI, active, dt, rho, K,
                                                                  GfWork{work, nwork_dev});
    return gf_launch_status();
}

extern "C" int gf_kkt_rhs_full(int B, int n, int m, int ld, int form, const int32_t* perm, const int32_t* nI,
                               const uint8_t* active, const double* F, const double* dt, const double* rho,
                               double* rhs, const int32_t* work, const int32_t* nwork_dev, int nwork, void* stream) {
    if (B <= 0 || n <= 0 || m < 0 || ld < n + m || !perm || !nI || !active || !F || !dt || !rho || !rhs) return GF_ERR_ARG;
    if (form != GF_FORM_ASYMMETRIC && form != GF_FORM_EXTENDED && form != GF_FORM_STANDARD) return GF_ERR_UNSUPPORTED;
    if (nwork <= 0) return GF_OK;
    kkt_full_rhs_kernel<<<nwork, pick_threads(n + m), 0, (cudaStream_t)stream>>>(n, m, ld, form, perm, nI, active, F, dt,
                                                                                 rho, rhs, GfWork{work, nwork_dev});
    return gf_launch_status();
}

extern "C" int gf_step_finish(int B, int n, int m, int ld, const double* xbase, const double* ybase, const double* sol,
                              const int32_t* perm, const int32_t* nI, const double* F, const double* dt,
                              const double* rho, const double* lb, const double* ub, double* xn, double* yn,
                              double* dx, double* dy, double* diff, const int32_t* work, const int32_t* nwork_dev,
                              int nwork, void* stream) {
    if (B <= 0 || n <= 0 || m < 0 || !xbase || !sol || !perm || !nI || !F || !dt || !rho || !lb || !ub || !xn || !diff)
        return GF_ERR_ARG;
    if (m > 0 && (!ybase || !yn)) return GF_ERR_ARG;
    if (nwork <= 0) return GF_OK;
    step_finish_kernel<<<nwork, pick_threads(n), 0, (cudaStream_t)stream>>>(
        n, m, ld, xbase, ybase, sol, perm, nI, F, dt, rho, lb, ub, xn, yn, dx, dy, diff, GfWork{work, nwork_dev});
    return gf_launch_status();
}

extern "C" int gf_check_terminate(int B, int n, int m, const double* x, const double* grad, const double* cons,
                                  const double* jty, const double* jtc, const double* obj, const double* lb,
                                  const double* ub, double opt_tol, double active_tol, double local_infeas_tol,
                                  double obj_lower_limit, int iteration_limit, const int32_t* iters, int32_t* status,
                                  double* total_res, const int32_t* work, const int32_t* nwork_dev, int nwork,
                                  void* stream) {
    if (B <= 0 || n <= 0 || m < 0 || !x || !grad || !obj || !lb || !ub || !iters || !status) return GF_ERR_ARG;
    if (m > 0 && (!cons || !jty || !jtc)) return GF_ERR_ARG;
    if (nwork <= 0) return GF_OK;
    terminate_kernel<<<nwork, pick_threads(n), 0, (cudaStream_t)stream>>>(
        n, m, x, grad, cons, jty, jtc, obj, lb, ub, opt_tol, active_tol, local_infeas_tol, obj_lower_limit,
        iteration_limit, iters, status, total_res, GfWork{work, nwork_dev});
    return gf_launch_status();
}

extern "C" int gf_dr_first(int B, const int32_t* status, const int32_t* info, const double* dt, const double* mid_norm,
                           const double* diff1, double newton_tol, double lamb_red, double lamb_min, int32_t* phase,
                           double* lamb_next, void* stream) {
    if (B <= 0 || !status || !dt || !mid_norm || !diff1 || !phase || !lamb_next) return GF_ERR_ARG;
    dr_first_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(B, status, info, dt, mid_norm, diff1, newton_tol,
                                                                       lamb_red, lamb_min, phase, lamb_next);
    return gf_launch_status();
}

extern "C" int gf_dr_second(int B, const double* dt, const double* diff1, const double* diff2, double theta_max,
                            double log_theta_ref, double K_P, double K_I, double lamb_min, double lamb_inc,
                            double* err_sum, int32_t* phase, double* lamb_next, double* theta, void* stream) {
    if (B <= 0 || !dt || !diff1 || !diff2 || !err_sum || !phase || !lamb_next) return GF_ERR_ARG;
    dr_second_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
        B, dt, diff1, diff2, theta_max, log_theta_ref, K_P, K_I, lamb_min, lamb_inc, err_sum, phase, lamb_next, theta);
    return gf_launch_status();
}

// Newton steps of one DistanceRatio outer iteration, counted from the phases after dr_second: one step for an instance that
// accepted at mid / final or rejected (phase 2..4), a second one where the second step was taken (phase 3, 4).
__global__ void count_newton_steps_kernel(int B, const int32_t* __restrict__ phase, unsigned long long* __restrict__ nsteps) {
    __shared__ int part[32];
    int c = 0;
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        const int ph = phase[b];
        c += (ph >= 2 && ph <= 4 ? 1 : 0) + (ph == 3 || ph == 4 ? 1 : 0);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += part[w];
        if (t > 0) atomicAdd(nsteps, (unsigned long long)t);
    }
}

extern "C" int gf_count_newton_steps(int B, const int32_t* phase, int64_t* nsteps, void* stream) {
    if (B <= 0 || !phase || !nsteps) return GF_ERR_ARG;
    count_newton_steps_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(B, phase, reinterpret_cast<unsigned long long*>(nsteps));
    return gf_launch_status();
}

extern "C" int gf_single_control(int B, int fixed, const int32_t* status, const int32_t* info, const double* dt,
                                 const double* mid_norm, const double* orig_norm, double newton_tol, double theta_max,
                                 double log_theta_ref, double K_P, double K_I, double lamb_red, double lamb_min,
                                 double lamb_inc, double lamb_init, double* err_sum, int32_t* phase, double* lamb_next,
                                 double* theta, int64_t* nsteps, void* stream) {
    if (B <= 0 || !status || !info || !dt || !mid_norm || !err_sum || !phase || !lamb_next || !theta || !nsteps ||
        (!fixed && !orig_norm))
        return GF_ERR_ARG;
    single_control_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
        B, fixed, status, info, dt, mid_norm, orig_norm, newton_tol, theta_max, log_theta_ref, K_P, K_I, lamb_red,
        lamb_min, lamb_inc, lamb_init, err_sum, phase, lamb_next, theta, reinterpret_cast<unsigned long long*>(nsteps));
    return gf_launch_status();
}

extern "C" int gf_exact_control(int B, int mode, int it, int last, const int32_t* status, const int32_t* info,
                                const double* dt, const double* val, const double* orig_norm, double newton_tol,
                                double rate_bound, double* curr, int32_t* live, int32_t* phase, double* lamb_next,
                                int64_t* nsteps, void* stream) {
    if (B <= 0 || mode < 0 || mode > 2 || !status || !info || !dt || !curr || !live || !phase || !lamb_next || !nsteps)
        return GF_ERR_ARG;
    if ((mode == 0 && !orig_norm) || (mode == 1 && !val)) return GF_ERR_ARG;
    exact_control_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
        B, mode, it, last, status, info, dt, val, orig_norm, newton_tol, rate_bound, curr, live, phase, lamb_next,
        reinterpret_cast<unsigned long long*>(nsteps));
    return gf_launch_status();
}

extern "C" int gf_commit(int B, int n, int m, const int32_t* phase, const double* lamb_next, double lamb_max,
                         int dual_norm_update, const double* xm, const double* ym, const double* gm, const double* cm,
                         const double* om, const double* xf, const double* yf, const double* gf, const double* cf,
                         const double* of, double* x, double* y, double* grad, double* cons, double* obj, double* lamb,
                         double* rho, int32_t* iters, int32_t* accepted, int32_t* status, void* stream) {
    if (B <= 0 || n <= 0 || m < 0 || !phase || !lamb_next || !xm || !gm || !om || !xf || !gf || !of || !x || !grad ||
        !obj || !lamb || !rho || !iters || !accepted || !status)
        return GF_ERR_ARG;
    commit_kernel<<<B, pick_threads(n), 0, (cudaStream_t)stream>>>(n, m, phase, lamb_next, lamb_max, dual_norm_update,
                                                                    xm, ym, gm, cm, om, xf, yf, gf, cf, of, x, y, grad,
                                                                    cons, obj, lamb, rho, iters, accepted, status);
    return gf_launch_status();
}

extern "C" int gf_build_worklist(int B, const int32_t* key, int lo, int hi, int invert, const int32_t* parent,
                                 const int32_t* parent_count, int32_t* list, int32_t* count, void* stream) {
    if (B <= 0 || !key || !list || !count || (parent != nullptr && parent_count == nullptr)) return GF_ERR_ARG;
    worklist_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(B, key, lo, hi, invert, parent, parent_count, list, count);
    return gf_launch_status();
}

extern "C" int gf_dt_from_lamb(int B, const double* lamb, double* dt, void* stream) {
    if (B <= 0 || !lamb || !dt) return GF_ERR_ARG;
    dt_from_lamb_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(B, lamb, dt);
    return gf_launch_status();
}
