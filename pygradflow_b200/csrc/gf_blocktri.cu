// Stage-structured KKT systems (cfg4: discretised optimal control): block-bidiagonal constraint Jacobian in a compact
// stage layout, diagonal Hessian of the Lagrangian.  Replaces, for such families,
//   symmetric_step_solver.py:27-94 (K = [[H_II + lamb I, J_I'], [J_I, -delta I]], rhs) + lu_solver.py:9-21
// by the Schur complement on the multipliers
//   M = delta I + J_I (H_II + lamb I)^-1 J_I'      (symmetric positive definite, block tridiagonal, S blocks of NX x NX)
//   M sy = J_I D^-1 b1 - (b2t - J_A b0),   dx_I = D^-1 (b1 - J_I' sy),   dx_A = b0
// i.e. the elimination order "primal variables first" that the dense LDL' path uses too; M is factorised by block
// cyclic reduction (log2 S levels instead of S NX sequential pivots): one CTA per matrix, NX x NX blocks inverted by
// in-warp Gauss-Jordan (four blocks per warp, width-8 shuffles), block products on the FP64 tensor pipe (DMMA 8x8x4).
//
// Compact layouts (per instance): Jc [S, NX + w, NX], w = NX + nu, column-major inside a stage: Jc[j][c][r] =
// d c_{j,r} / d (column c), columns 0..NX-1 = x_j (variables (j-1) w + c; zero block for j = 0), columns NX..NX+w-1 = z_j
// (variables j w + c - NX) -- row sums (J v), column sums (J' v) and DMMA fragments all read it coalesced; Hd [n] the
// diagonal of the Hessian.  Factors per block i, eliminated at level s = lowbit(i) with neighbours i - s and i + s:
// Tinv = inv(T_i), P = Tinv K[i, i-s], Q = Tinv K[i, i+s]   ([S, NX*NX] each; block 0 is the last one, P = Q = 0).
#include "gf_common.cuh"
#include "../../include/gradflow_b200.h"

#ifdef GF_STAGE_TRACE
__device__ unsigned long long g_stage_trace[32];
#define ST_DECL long long st_t_ = clock64()
#define ST_MARK(i) do { if (threadIdx.x == 0) { long long n_ = clock64(); atomicAdd(&g_stage_trace[i], (unsigned long long)(n_ - st_t_)); st_t_ = n_; } } while (0)
#else
#define ST_DECL
#define ST_MARK(i)
#endif

namespace {

constexpr int NX = 8;          // states per stage = block size of M
constexpr int BS = NX * NX;    // doubles per block

// ---- compact evaluators of the OCP family (same arithmetic as ocp_jac_kernel / ocp_hess_kernel of gf_eval.cu) ----
// diag_only != 0: the buffer already holds the Jacobian of an earlier point; only the entries that depend on z -- the
// diagonal of d c_j / d x_j, through cos x_j -- are rewritten.
__global__ void ocp_jac_banded_kernel(int S, int nx, int nu, double h, const double* __restrict__ A,
                                      const double* __restrict__ Bm, const double* __restrict__ z,
                                      double* __restrict__ Jc, int diag_only, GfWork work) {
    const int b = gf_instance(work, blockIdx.x);
    if (b < 0) return;
    const int w = nx + nu, n = S * w, jw = nx + w;
    const double* zb = z + (size_t)b * n;
    const double* Ab = A + (size_t)b * S * nx * nx;
    const double* Bb = Bm + (size_t)b * S * nx * nu;
    double* Jb = Jc + (size_t)b * S * nx * jw;
    if (diag_only) {
        for (int e = threadIdx.x; e < S * nx; e += blockDim.x) {
            const int j = e / nx, r = e - j * nx;
            if (j == 0) continue;
            const double xe = zb[(size_t)(j - 1) * w + r];
            const double ee = __dmul_rn(h, __dadd_rn(Ab[((size_t)j * nx + r) * nx + r], __dmul_rn(1.0, __dmul_rn(0.1, cos(xe)))));
            Jb[((size_t)j * jw + r) * nx + r] = -__dadd_rn(1.0, ee);
        }
        return;
    }
    for (int e = threadIdx.x; e < S * nx * jw; e += blockDim.x) {
        const int j = e / (nx * jw), t = e - j * nx * jw, c = t / nx, r = t - c * nx;
        double v;
        if (c < nx) {  // d c_j / d x_j
            if (j >= 1) {
                const double d = (r == c) ? 1.0 : 0.0;
                double cs = 0.0;
                if (r == c) cs = __dmul_rn(0.1, cos(zb[(size_t)(j - 1) * w + c]));
                const double ee = __dmul_rn(h, __dadd_rn(Ab[((size_t)j * nx + r) * nx + c], __dmul_rn(d, cs)));
                v = -__dadd_rn(d, ee);
            } else {
                v = 0.0;
            }
        } else if (c < 2 * nx) {  // d c_j / d x_{j+1} = I
            v = (r == c - nx) ? 1.0 : 0.0;
        } else {                  // d c_j / d u_j = -h B_j
            v = -__dmul_rn(h, Bb[((size_t)j * nx + r) * nu + (c - 2 * nx)]);
        }
        Jb[e] = v;
    }
}

__global__ void ocp_hess_diag_kernel(int S, int nx, int nu, double c1, const double* __restrict__ Q,
                                     const double* __restrict__ R, const double* __restrict__ z,
                                     const double* __restrict__ y, double* __restrict__ Hd, GfWork work) {
    const int b = gf_instance(work, blockIdx.x);
    if (b < 0) return;
    const int w = nx + nu, n = S * w, m = S * nx;
    const double* zb = z + (size_t)b * n;
    const double* yb = y + (size_t)b * m;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int j = i / w, c = i - j * w;
        double v;
        if (c < nx) {
            v = Q[((size_t)b * S + j) * nx + c];
            if (j + 1 < S) v = __dadd_rn(v, __dmul_rn(__dmul_rn(yb[(size_t)(j + 1) * nx + c], c1), sin(zb[i])));
        } else {
            v = R[((size_t)b * S + j) * nu + (c - nx)];
        }
        Hd[(size_t)b * n + i] = v;
    }
}

// (J' v)_i for variable i = j w + k: column NX + k of stage j and, for a state, column k of stage j + 1; the NX entries of
// a column are contiguous (two 32-byte sectors).
__device__ __forceinline__ void jt_gather3(const double* __restrict__ Jb, int S, int w, int jw, int i,
                                           const double* v0, const double* v1, const double* v2, double& o0, double& o1,
                                           double& o2) {
    const int j = i / w, k = i - j * w;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
    const double* col = Jb + ((size_t)j * jw + NX + k) * NX;
#pragma unroll
    for (int r = 0; r < NX; r += 2) {
        const double2 e = *reinterpret_cast<const double2*>(col + r);
        a0 += e.x * v0[j * NX + r] + e.y * v0[j * NX + r + 1];
        if (v1) a1 += e.x * v1[j * NX + r] + e.y * v1[j * NX + r + 1];
        if (v2) a2 += e.x * v2[j * NX + r] + e.y * v2[j * NX + r + 1];
    }
    if (k < NX && j + 1 < S) {
        const double* colp = Jb + ((size_t)(j + 1) * jw + k) * NX;
#pragma unroll
        for (int r = 0; r < NX; r += 2) {
            const double2 e = *reinterpret_cast<const double2*>(colp + r);
            a0 += e.x * v0[(j + 1) * NX + r] + e.y * v0[(j + 1) * NX + r + 1];
            if (v1) a1 += e.x * v1[(j + 1) * NX + r] + e.y * v1[(j + 1) * NX + r + 1];
            if (v2) a2 += e.x * v2[(j + 1) * NX + r] + e.y * v2[(j + 1) * NX + r + 1];
        }
    }
    o0 = a0; o1 = a1; o2 = a2;
}

// Iterate.aug_lag_deriv_x (iterate.py:91-94) with the compact Jacobian: dL = grad + J'(rho c + y), jty = J'y, jtc = J'c.
__global__ void stage_aug_lag_grad_kernel(int S, int nu, const double* __restrict__ Jc, const double* __restrict__ grad,
                                          const double* __restrict__ cons, const double* __restrict__ y,
                                          const double* __restrict__ rho, double* __restrict__ dL,
                                          double* __restrict__ jty, double* __restrict__ jtc, GfWork work) {
    const int b = gf_instance(work, blockIdx.x);
    if (b < 0) return;
    const int w = NX + nu, n = S * w, m = S * NX, jw = NX + w;
    extern __shared__ double sm[];
    double* sv = sm;          // rho c + y
    double* sy = sm + m;      // y
    double* sc = sm + 2 * m;  // c
    const double rb = rho[b];
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
        const double c = cons[(size_t)b * m + i], yy = y[(size_t)b * m + i];
        sv[i] = __dadd_rn(__dmul_rn(rb, c), yy);
        sy[i] = yy;
        sc[i] = c;
    }
    __syncthreads();
    const double* Jb = Jc + (size_t)b * m * jw;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double a0, a1, a2;
        jt_gather3(Jb, S, w, jw, i, sv, jty ? sy : nullptr, jtc ? sc : nullptr, a0, a1, a2);
        if (dL) dL[(size_t)b * n + i] = __dadd_rn(grad[(size_t)b * n + i], a0);
        if (jty) jty[(size_t)b * n + i] = a1;
        if (jtc) jtc[(size_t)b * n + i] = a2;
    }
}

// ---- 8x8 block primitives (row-major blocks of 64 doubles) ---------------------------------------------------------
// D (C-fragment registers) += sign * op(A) op(B); TA / TB: use the transpose of the stored block.
template <bool TA, bool TB, bool NEG>
__device__ __forceinline__ void blk_mma(double& c0, double& c1, const double* A, const double* B) {
    const int t = threadIdx.x & 31, g = t >> 2, q = t & 3;
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const int kk = q + 4 * h;
        double a = TA ? A[kk * NX + g] : A[g * NX + kk];        // a = opA[g][kk]
        const double bv = TB ? B[g * NX + kk] : B[kk * NX + g];  // b = opB[kk][g]
        if (NEG) a = -a;
        dmma884(c0, c1, a, bv);
    }
}
// C-fragment <-> memory: lane t owns elements 2t, 2t + 1 of the row-major block (a coalesced 512-byte access)
__device__ __forceinline__ void blk_load_c(double& c0, double& c1, const double* C) {
    const double2 v = *reinterpret_cast<const double2*>(C + 2 * (threadIdx.x & 31));
    c0 = v.x; c1 = v.y;
}
__device__ __forceinline__ void blk_store_c(double c0, double c1, double* C) {
    *reinterpret_cast<double2*>(C + 2 * (threadIdx.x & 31)) = make_double2(c0, c1);
}

// 1 / d to ~1 ulp: MUFU seed + two Newton steps (the reciprocal sits on the pivot-to-pivot chain)
__device__ __forceinline__ double fast_rcp(double d) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    double e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    return r;
}

// In-place inverse of an SPD 8x8 block by Gauss-Jordan without pivoting; the 8 lanes of a width-8 group hold one row
// each.  Returns false when a pivot is not positive / not finite.
__device__ __forceinline__ bool gj_inverse8(double (&a)[NX], int r) {
    bool ok = true;
#pragma unroll
    for (int k = 0; k < NX; k++) {
        double rk[NX];
#pragma unroll
        for (int c = 0; c < NX; c++) rk[c] = __shfl_sync(0xffffffffu, a[c], k, NX);
        const double p = rk[k];
        if (!(p > 0.0) || !(p < 1.0e300)) ok = false;
        const double pinv = fast_rcp(p);
        const double f = (r == k) ? 0.0 : a[k] * pinv;    // row k itself: scaled below
#pragma unroll
        for (int c = 0; c < NX; c++) {
            if (c == k) a[c] = (r == k) ? pinv : -f;
            else a[c] = (r == k) ? a[c] * pinv : fma(-f, rk[c], a[c]);
        }
    }
    return ok;
}

// ---- factorisation ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) stage_factor_kernel(int S, int nu, const double* __restrict__ Jc,
                                                           const double* __restrict__ Hd,
                                                           const uint8_t* __restrict__ active,
                                                           const double* __restrict__ dt, const double* __restrict__ rho,
                                                           double* __restrict__ Tinv, double* __restrict__ Pf,
                                                           double* __restrict__ Qf, int32_t* __restrict__ info,
                                                           int32_t* __restrict__ nneg, GfWork work) {
    const int b = gf_instance(work, blockIdx.x);
    if (b < 0) return;
    const int w = NX + nu, n = S * w, m = S * NX, jw = NX + w;
    extern __shared__ double sm[];
    double* T = sm;                    // [S][64]  diagonal blocks; the inverse once a block is eliminated
    double* C = T + (size_t)S * BS;    // [S][64]  coupling of block i to its current left neighbour, K[i, i - s]
    double* invD = C + (size_t)S * BS; // [n]
    const int nwarp = blockDim.x >> 5, wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* X = invD + n + (size_t)wid * BS;  // per-warp 8x8 scratch
    __shared__ int s_bad;
    if (threadIdx.x == 0) s_bad = 0;
    const double lamb = 1.0 / dt[b];
    const double delta = lamb / (1.0 + lamb * rho[b]);
    ST_DECL;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const bool act = active[(size_t)b * n + i] != 0;
        const double d = Hd[(size_t)b * n + i] + lamb;
        if (!act && !(d > 0.0)) atomicExch(&s_bad, !(d == d) ? GF_INFO_NONFINITE : GF_INFO_NOT_QUASIDEFINITE);
        invD[i] = act ? 0.0 : 1.0 / d;
    }
    __syncthreads();
    ST_MARK(0);
    // ---- M blocks by DMMA straight from the compact Jacobian: with A = J D^-1 (8 x K) and B = J' (K x 8) the a- and
    // b-fragments of a k-step are the SAME element J[g][kk] (scaled for a), one coalesced load each.
    const double* Jb = Jc + (size_t)b * m * jw;
    {
        const int g = lane >> 2, q = lane & 3;
        for (int j = wid; j < S; j += nwarp) {
            const double* Js = Jb + (size_t)j * jw * NX;         // stage j: [jw][NX]
            double t0 = (g == 2 * q) ? delta : 0.0, t1 = (g == 2 * q + 1) ? delta : 0.0, c0 = 0.0, c1 = 0.0;
            double v[8], vp[2];
            const int nk = jw / 4;  // k-steps of four columns (jw = 2 NX + nu; nu a multiple of 4 is required)
#pragma unroll
            for (int ks = 0; ks < 8; ks++)
                if (ks < nk) v[ks] = Js[(size_t)(4 * ks + q) * NX + g];
            if (j >= 1) {
#pragma unroll
                for (int ks = 0; ks < 2; ks++) vp[ks] = Jb[((size_t)(j - 1) * jw + NX + 4 * ks + q) * NX + g];
            }
#pragma unroll
            for (int ks = 0; ks < 8; ks++) {
                if (ks < nk) {
                    const int col = 4 * ks + q;
                    double sc;
                    if (col < NX) sc = j >= 1 ? invD[(j - 1) * w + col] : 0.0;
                    else sc = invD[j * w + col - NX];
                    const double a = v[ks] * sc;
                    dmma884(t0, t1, a, v[ks]);
                    if (ks < 2 && j >= 1) dmma884(c0, c1, a, vp[ks]);   // M_{j,j-1} = (Jp_j D^-1) Jz_{j-1}[:, x part]'
                }
            }
            blk_store_c(t0, t1, T + (size_t)j * BS);
            blk_store_c(c0, c1, C + (size_t)j * BS);
        }
    }
    __syncthreads();
    ST_MARK(1);
    double* Tg = Tinv + (size_t)b * S * BS;
    double* Pg = Pf + (size_t)b * S * BS;
    double* Qg = Qf + (size_t)b * S * BS;
    const int grp = lane >> 3, row = lane & 7;
    // ---- block cyclic reduction
    for (int s = 1; s < S; s <<= 1) {
        // (a) eliminated blocks i = s, 3s, 5s, ...: T_i <- inv(T_i); four blocks per warp
        const int nel = (S - s + 2 * s - 1) / (2 * s);
        for (int q0 = wid * 4; q0 < nel; q0 += nwarp * 4) {
            const int q = q0 + grp;
            const int i = s + 2 * s * q;
            const bool have = q < nel;
            double a[NX];
#pragma unroll
            for (int c = 0; c < NX; c++) a[c] = have ? T[(size_t)i * BS + row * NX + c] : (c == row ? 1.0 : 0.0);
            const bool ok = gj_inverse8(a, row);
            if (have) {
                if (!ok) atomicExch(&s_bad, i * NX + 1);
#pragma unroll
                for (int c = 0; c < NX; c++) T[(size_t)i * BS + row * NX + c] = a[c];
            }
        }
        __syncthreads();
        ST_MARK(2);
        // (b) remaining blocks k = 0, 2s, 4s, ...: Schur updates from the eliminated neighbours k - s and k + s; the
        // products Tinv K[i, .] they need are the factors P / Q of those blocks and go to global memory from here
        const int nrem = (S + 2 * s - 1) / (2 * s);
        for (int q = wid; q < nrem; q += nwarp) {
            const int k = 2 * s * q;
            double t0, t1, n0 = 0.0, n1 = 0.0;
            blk_load_c(t0, t1, T + (size_t)k * BS);
            const double* Ck = C + (size_t)k * BS;
            if (k >= s) {
                const double* Ti = T + (size_t)(k - s) * BS;  // inverse
                double x0 = 0.0, x1 = 0.0;
                blk_mma<false, true, false>(x0, x1, Ti, Ck);   // X = Tinv C_k' = Q_{k-s}
                blk_store_c(x0, x1, X);
                blk_store_c(x0, x1, Qg + (size_t)(k - s) * BS);
                __syncwarp();
                blk_mma<false, false, true>(t0, t1, Ck, X);    // T_k -= C_k X
                __syncwarp();
                if (k >= 2 * s) {
                    x0 = 0.0; x1 = 0.0;
                    blk_mma<false, false, false>(x0, x1, Ti, C + (size_t)(k - s) * BS);  // X = Tinv C_{k-s} = P_{k-s}
                    blk_store_c(x0, x1, X);
                    __syncwarp();
                    blk_mma<false, false, true>(n0, n1, Ck, X);  // C_k(new) = -C_k X
                    __syncwarp();
                }
            }
            if (k + s < S) {
                const double* Ti = T + (size_t)(k + s) * BS;
                const double* Ci = C + (size_t)(k + s) * BS;
                double x0 = 0.0, x1 = 0.0;
                blk_mma<false, false, false>(x0, x1, Ti, Ci);  // X = Tinv C_{k+s} = P_{k+s}
                blk_store_c(x0, x1, X);
                blk_store_c(x0, x1, Pg + (size_t)(k + s) * BS);
                if (k + 2 * s >= S) blk_store_c(0.0, 0.0, Qg + (size_t)(k + s) * BS);  // no right neighbour
                __syncwarp();
                blk_mma<true, false, true>(t0, t1, Ci, X);      // T_k -= C_{k+s}' X
                __syncwarp();
            }
            blk_store_c(t0, t1, T + (size_t)k * BS);
            blk_store_c(n0, n1, C + (size_t)k * BS);
        }
        __syncthreads();
        ST_MARK(3);
    }
    // ---- the last block
    if (wid == 0) {
        double a[NX];
        const bool have = grp == 0;
#pragma unroll
        for (int c = 0; c < NX; c++) a[c] = have ? T[row * NX + c] : (c == row ? 1.0 : 0.0);
        const bool ok = gj_inverse8(a, row);
        if (have) {
            if (!ok) atomicExch(&s_bad, 1);
#pragma unroll
            for (int c = 0; c < NX; c++) T[row * NX + c] = a[c];
        }
    }
    __syncthreads();
    // all inverses, coalesced; block 0 has no neighbours
    for (int e = threadIdx.x; e < S * BS; e += blockDim.x) Tg[e] = T[e];
    for (int e = threadIdx.x; e < BS; e += blockDim.x) { Pg[e] = 0.0; Qg[e] = 0.0; }
    ST_MARK(4);
    if (threadIdx.x == 0) {
        info[b] = s_bad;
        nneg[b] = m;  // M positive definite <=> K has exactly m negative eigenvalues
    }
}

// ---- solve: sol = (dx in the natural order with dx_A = b0, sy) for the scaled residual F --------------------------
__device__ __forceinline__ double row_dot8(const double* __restrict__ Mrow, const double* v) {
    const double4 a = *reinterpret_cast<const double4*>(Mrow), c = *reinterpret_cast<const double4*>(Mrow + 4);
    return a.x * v[0] + a.y * v[1] + a.z * v[2] + a.w * v[3] + c.x * v[4] + c.y * v[5] + c.z * v[6] + c.w * v[7];
}
__device__ __forceinline__ double col_dot8(const double* __restrict__ M, int r, const double* v) {
    double acc = 0.0;
#pragma unroll
    for (int c = 0; c < NX; c++) acc += M[c * NX + r] * v[c];
    return acc;
}

__global__ void __launch_bounds__(256) stage_solve_kernel(int S, int nu, const double* __restrict__ Jc,
                                                          const double* __restrict__ Hd,
                                                          const uint8_t* __restrict__ active,
                                                          const double* __restrict__ F, const double* __restrict__ dt,
                                                          const double* __restrict__ rho, const double* __restrict__ Tinv,
                                                          const double* __restrict__ Pf, const double* __restrict__ Qf,
                                                          double* __restrict__ sol, int ldsol, GfWork work) {
    const int b = gf_instance(work, blockIdx.x);
    if (b < 0) return;
    const int w = NX + nu, n = S * w, m = S * NX, jw = NX + w;
    extern __shared__ double sm[];
    double* wv = sm;        // [n]  w_i = b0 (active) or D^-1 b1 (inactive); later dx
    double* rr = wv + n;    // [m]  right-hand side of M / solution sy
    const double dtb = dt[b];
    const double lamb = 1.0 / dtb;
    const double fact = 1.0 / (1.0 + lamb * rho[b]);
    const double* Fb = F + (size_t)b * (n + m);
    const double* Jb = Jc + (size_t)b * m * jw;
    ST_DECL;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const bool act = active[(size_t)b * n + i] != 0;
        const double fx = Fb[i];
        wv[i] = act ? __dmul_rn(dtb, fx) : fx / (Hd[(size_t)b * n + i] + lamb);   // scaled_step_solver.py:56-57
    }
    __syncthreads();
    for (int i = threadIdx.x; i < m; i += blockDim.x) {   // J w - fact F_y; the 8 rows of a stage read each column coalesced
        const int j = i / NX, r = i - j * NX;
        const double* Js = Jb + (size_t)j * jw * NX + r;
        double acc = 0.0;
        for (int k = 0; k < w; k++) acc += Js[(size_t)(NX + k) * NX] * wv[j * w + k];
        if (j >= 1) {
#pragma unroll
            for (int k = 0; k < NX; k++) acc += Js[(size_t)k * NX] * wv[(j - 1) * w + k];
        }
        rr[i] = acc - fact * Fb[n + i];
    }
    __syncthreads();
    ST_MARK(8);
    const double* Tg = Tinv + (size_t)b * S * BS;
    const double* Pg = Pf + (size_t)b * S * BS;
    const double* Qg = Qf + (size_t)b * S * BS;
    const int r = threadIdx.x & 7, slot = threadIdx.x >> 3, nslot = blockDim.x >> 3;
    int top = 1;
    while (top * 2 < S) top *= 2;
    // forward elimination: r_k -= K[k, k-s] Tinv_{k-s} r_{k-s} + K[k, k+s] Tinv_{k+s} r_{k+s} = Q_{k-s}' r_{k-s} + P_{k+s}' r_{k+s}
    for (int s = 1; s < S; s <<= 1) {
        const int nrem = (S + 2 * s - 1) / (2 * s);
        for (int q = slot; q < nrem; q += nslot) {
            const int k = 2 * s * q;
            double acc = 0.0;
            if (k >= s) acc += col_dot8(Qg + (size_t)(k - s) * BS, r, rr + (k - s) * NX);
            if (k + s < S) acc += col_dot8(Pg + (size_t)(k + s) * BS, r, rr + (k + s) * NX);
            rr[k * NX + r] -= acc;
        }
        __syncthreads();
    }
    ST_MARK(9);
    // back substitution: v_0 = Tinv_0 r_0;  v_i = Tinv_i r_i - P_i v_{i-s} - Q_i v_{i+s}
    {
        double v0 = 0.0;
        if (slot == 0) v0 = row_dot8(Tg + r * NX, rr);
        __syncthreads();
        if (slot == 0) rr[r] = v0;
        __syncthreads();
    }
    for (int s = top; s >= 1; s >>= 1) {
        const int nel = (S - s + 2 * s - 1) / (2 * s);
        double vals[4];
        int cnt = 0;
        for (int q = slot; q < nel; q += nslot) {
            const int i = s + 2 * s * q;
            double acc = row_dot8(Tg + (size_t)i * BS + r * NX, rr + i * NX) - row_dot8(Pg + (size_t)i * BS + r * NX, rr + (i - s) * NX);
            if (i + s < S) acc -= row_dot8(Qg + (size_t)i * BS + r * NX, rr + (i + s) * NX);
            vals[cnt++ & 3] = acc;
        }
        __syncthreads();   // every row of r_i has been read before it is overwritten by v_i
        cnt = 0;
        for (int q = slot; q < nel; q += nslot) rr[(s + 2 * s * q) * NX + r] = vals[cnt++ & 3];
        __syncthreads();
    }
    ST_MARK(10);
    // dx_I = D^-1 (b1 - J' sy) = w - D^-1 J' sy; dx_A = b0
    double* out = sol + (size_t)b * ldsol;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double v = wv[i];
        if (active[(size_t)b * n + i] == 0) {
            double a0, a1, a2;
            jt_gather3(Jb, S, w, jw, i, rr, nullptr, nullptr, a0, a1, a2);
            v -= a0 / (Hd[(size_t)b * n + i] + lamb);
        }
        out[i] = v;
    }
    for (int i = threadIdx.x; i < m; i += blockDim.x) out[n + i] = rr[i];
    for (int i = n + m + threadIdx.x; i < ldsol; i += blockDim.x) out[i] = 0.0;
    ST_MARK(11);
}

}  // namespace

extern "C" int gf_ocp_jac_banded(int B, int S, int nx, int nu, double h, const double* A, const double* Bm,
                                 const double* z, double* Jc, int diag_only, const int32_t* work,
                                 const int32_t* nwork_dev, int nwork, void* stream) {
    if (B <= 0 || S <= 0 || nx <= 0 || nu <= 0 || !A || !Bm || !z || !Jc) return GF_ERR_ARG;
    if (nwork <= 0) return GF_OK;
    ocp_jac_banded_kernel<<<nwork, diag_only ? 256 : 512, 0, (cudaStream_t)stream>>>(S, nx, nu, h, A, Bm, z, Jc, diag_only,
                                                                                      GfWork{work, nwork_dev});
    return gf_launch_status();
}

extern "C" int gf_ocp_hess_diag(int B, int S, int nx, int nu, double c1, const double* Q, const double* R,
                                const double* z, const double* y, double* Hd, const int32_t* work,
                                const int32_t* nwork_dev, int nwork, void* stream) {
    if (B <= 0 || S <= 0 || nx <= 0 || nu <= 0 || !Q || !R || !z || !y || !Hd) return GF_ERR_ARG;
    if (nwork <= 0) return GF_OK;
    ocp_hess_diag_kernel<<<nwork, 512, 0, (cudaStream_t)stream>>>(S, nx, nu, c1, Q, R, z, y, Hd, GfWork{work, nwork_dev});
    return gf_launch_status();
}

extern "C" int gf_stage_aug_lag_grad(int B, int S, int nx, int nu, const double* Jc, const double* grad,
                                     const double* cons, const double* y, const double* rho, double* dL, double* jty,
                                     double* jtc, const int32_t* work, const int32_t* nwork_dev, int nwork,
                                     void* stream) {
    if (B <= 0 || S <= 0 || nu <= 0 || !Jc || !grad || !cons || !y || !rho) return GF_ERR_ARG;
    if (nx != NX) return GF_ERR_UNSUPPORTED;
    if (nwork <= 0) return GF_OK;
    const size_t smem = (size_t)3 * S * NX * sizeof(double);
    if (smem > 200 * 1024) return GF_ERR_UNSUPPORTED;
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(stage_aug_lag_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    stage_aug_lag_grad_kernel<<<nwork, 512, smem, (cudaStream_t)stream>>>(S, nu, Jc, grad, cons, y, rho, dL, jty, jtc,
                                                                           GfWork{work, nwork_dev});
    return gf_launch_status();
}

extern "C" int gf_stage_kkt_factor(int B, int S, int nx, int nu, const double* Jc, const double* Hd,
                                   const uint8_t* active, const double* dt, const double* rho, double* Tinv, double* Lc,
                                   double* Uc, int32_t* info, int32_t* nneg, const int32_t* work,
                                   const int32_t* nwork_dev, int nwork, void* stream) {
    if (B <= 0 || S <= 0 || nu <= 0 || !Jc || !Hd || !active || !dt || !rho || !Tinv || !Lc || !Uc || !info || !nneg)
        return GF_ERR_ARG;
    if (nx != NX || (nu % 4) != 0 || nu > 16) return GF_ERR_UNSUPPORTED;
    if (nwork <= 0) return GF_OK;
    const int w = NX + nu, threads = 256;
    const size_t smem = ((size_t)2 * S * BS + (size_t)S * w + (size_t)(threads / 32) * BS) * sizeof(double);
    if (smem > 220 * 1024) return GF_ERR_UNSUPPORTED;
    cudaFuncSetAttribute(stage_factor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    stage_factor_kernel<<<nwork, threads, smem, (cudaStream_t)stream>>>(S, nu, Jc, Hd, active, dt, rho, Tinv, Lc, Uc, info,
                                                                         nneg, GfWork{work, nwork_dev});
    return gf_launch_status();
}

extern "C" int gf_stage_kkt_solve(int B, int S, int nx, int nu, const double* Jc, const double* Hd,
                                  const uint8_t* active, const double* F, const double* dt, const double* rho,
                                  const double* Tinv, const double* Lc, const double* Uc, double* sol, int ldsol,
                                  const int32_t* work, const int32_t* nwork_dev, int nwork, void* stream) {
    if (B <= 0 || S <= 0 || nu <= 0 || !Jc || !Hd || !active || !F || !dt || !rho || !Tinv || !Lc || !Uc || !sol)
        return GF_ERR_ARG;
    if (nx != NX) return GF_ERR_UNSUPPORTED;
    const int w = NX + nu, n = S * w, m = S * NX;
    if (ldsol < n + m) return GF_ERR_ARG;
    if (nwork <= 0) return GF_OK;
    const size_t smem = (size_t)(n + m) * sizeof(double);
    if (smem > 200 * 1024) return GF_ERR_UNSUPPORTED;
    if (smem > 48 * 1024) cudaFuncSetAttribute(stage_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    stage_solve_kernel<<<nwork, 256, smem, (cudaStream_t)stream>>>(S, nu, Jc, Hd, active, F, dt, rho, Tinv, Lc, Uc, sol,
                                                                    ldsol, GfWork{work, nwork_dev});
    return gf_launch_status();
}
