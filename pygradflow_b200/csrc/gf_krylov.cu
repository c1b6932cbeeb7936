// Batched iterative LinearSolvers: restarted GMRES and MINRES, one CTA per instance.
//
// Replaces pygradflow/linear_solver/gmres_solver.py:7-35 (scipy.sparse.linalg.gmres(mat, rhs, maxiter=n, x0, atol=1e-8),
// default rtol = 1e-5, restart = 20, plus the wrapper's own "residual of x0 already below atol" early return) and
// pygradflow/linear_solver/minres_solver.py:6-24 (scipy.sparse.linalg.minres(mat, rhs, x0), rtol = 1e-5, maxiter = 5 n).
// The iteration is scipy's (scipy/sparse/linalg/_isolve/iterative.py:gmres, minres.py:minres; SciPy 1.18, the version the
// oracle runs) statement by statement -- same recurrences, same stopping tests, same tolerance bookkeeping -- so an instance
// stops after the same number of matrix-vector products as the reference; only the summation order inside dot products
// differs (fixed block tree here, BLAS there).
//
// Layout: the matrix K[b] (ld x ld, row-major, order N_b in its top-left corner) stays in HBM and is streamed once per
// matrix-vector product -- the bound of both methods: 8 N^2 bytes per product.  A x: one warp per row, lanes across the
// columns (coalesced), the vector staged in shared memory; A' x: one thread per column, rows streamed (coalesced across
// the threads).  Krylov vectors live in a caller-provided scratch area (L1/L2 resident: a few rows of ld doubles per
// instance); every thread owns the same vector elements throughout, so element-wise updates need no barrier and only
// products and reductions synchronise the CTA.
#include "gf_common.cuh"
#include "../../include/gradflow_b200.h"

namespace {

constexpr int KR_THREADS = 512;
constexpr int KR_WARPS = KR_THREADS / 32;
constexpr double KR_EPS = 2.220446049250313e-16;  // np.finfo(float64).eps

struct KrMat {
    const double* A;
    int ld;
    int N;
    int trans;
};

// y[0..N) = op(A) xs, xs in shared memory.  All threads call; ends with a barrier so y is visible to the CTA.
__device__ __forceinline__ void kr_matvec(const KrMat& M, const double* __restrict__ xs, double* __restrict__ y) {
    const int N = M.N, ld = M.ld;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (!M.trans) {
        for (int r = wid; r < N; r += KR_WARPS) {
            const double* row = M.A + (size_t)r * ld;
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
            int c = lane;
            for (; c + 96 < N; c += 128) {
                const double v0 = __ldg(row + c), v1 = __ldg(row + c + 32), v2 = __ldg(row + c + 64),
                             v3 = __ldg(row + c + 96);
                a0 = fma(v0, xs[c], a0);
                a1 = fma(v1, xs[c + 32], a1);
                a2 = fma(v2, xs[c + 64], a2);
                a3 = fma(v3, xs[c + 96], a3);
            }
            for (; c < N; c += 32) a0 = fma(__ldg(row + c), xs[c], a0);
            const double s = warp_sum((a0 + a1) + (a2 + a3));
            if (lane == 0) y[r] = s;
        }
    } else {
        for (int c = threadIdx.x; c < N; c += KR_THREADS) {
            const double* col = M.A + c;
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
            int r = 0;
            for (; r + 3 < N; r += 4) {
                const double v0 = __ldg(col + (size_t)r * ld), v1 = __ldg(col + (size_t)(r + 1) * ld),
                             v2 = __ldg(col + (size_t)(r + 2) * ld), v3 = __ldg(col + (size_t)(r + 3) * ld);
                a0 = fma(v0, xs[r], a0);
                a1 = fma(v1, xs[r + 1], a1);
                a2 = fma(v2, xs[r + 2], a2);
                a3 = fma(v3, xs[r + 3], a3);
            }
            for (; r < N; ++r) a0 = fma(__ldg(col + (size_t)r * ld), xs[r], a0);
            y[c] = (a0 + a1) + (a2 + a3);
        }
    }
    __syncthreads();
}

__device__ __forceinline__ double kr_dot(const double* a, const double* b, int N, double* red) {
    double s = 0.0;
    for (int i = threadIdx.x; i < N; i += KR_THREADS) s = fma(a[i], b[i], s);
    return block_sum(s, red);
}

// LAPACK dlartg (3.10+, la_lartg.f90) on the unscaled branch, which is the one scipy's gmres reaches for the
// magnitudes a Hessenberg column can take here; the scaled branch is restated for completeness.
__device__ void kr_lartg(double f, double g, double& c, double& s, double& r) {
    const double safmin = 2.2250738585072014e-308, safmax = 1.0 / safmin;
    const double rtmin = 1.4916681462400413e-154;           // sqrt(safmin)
    const double rtmax = 4.7403759540545887e+153;           // sqrt(safmax / 2)
    const double f1 = fabs(f), g1 = fabs(g);
    if (g == 0.0) {
        c = 1.0; s = 0.0; r = f;
    } else if (f == 0.0) {
        c = 0.0; s = copysign(1.0, g); r = g1;
    } else if (f1 > rtmin && f1 < rtmax && g1 > rtmin && g1 < rtmax) {
        const double d = sqrt(f * f + g * g);
        c = f1 / d;
        r = copysign(d, f);
        s = g / r;
    } else {
        const double u = fmin(safmax, fmax(safmin, fmax(f1, g1)));
        const double fs = f / u, gs = g / u;
        const double d = sqrt(fs * fs + gs * gs);
        c = fabs(fs) / d;
        r = copysign(d, f);
        s = gs / r;
        r *= u;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// GMRES(restart): scratch per instance = (restart + 4) rows of ld doubles: V[0..restart], x, w, r.
__global__ void __launch_bounds__(KR_THREADS) gmres_kernel(int ld, int ldr, int Nmax, const int32_t* __restrict__ Nvec,
                                                           const double* __restrict__ Kall, double* __restrict__ rhs_all,
                                                           const double* __restrict__ x0_all,
                                                           const uint8_t* __restrict__ mask_all, int nmask, int trans,
                                                           int restart_in, double rtol, double atol_in,
                                                           double* __restrict__ scratch_all, int32_t* __restrict__ info,
                                                           int32_t* __restrict__ iters, GfWork work) {
    const int b = gf_instance(work, blockIdx.x);
    if (b < 0) return;
    const int N = Nvec != nullptr ? Nvec[b] : Nmax;
    const int tid = threadIdx.x;
    extern __shared__ double sm[];
    double* xs = sm;                   // Nmax: staged input of the matrix-vector product
    double* red = xs + Nmax;           // 32
    double* hh = red + 32;             // restart * (restart + 1): h[col][row] like scipy's h[col, k]
    double* giv = hh + restart_in * (restart_in + 1);  // 2 * restart
    double* S = giv + 2 * restart_in;  // restart + 1
    double* yv = S + restart_in + 1;   // restart + 1
    double* sc = yv + restart_in + 1;  // 4 broadcast scalars

    double* rhs = rhs_all + (size_t)b * ldr;
    if (N <= 0) {
        if (tid == 0) { info[b] = 0; if (iters) iters[b] = 0; }
        return;
    }
    const int restart = min(restart_in, N);
    const int maxiter = N;  // gmres_solver.py:27 maxiter = n
    double* base = scratch_all + (size_t)b * (size_t)(restart_in + 4) * ld;
    double* V = base;
    double* x = base + (size_t)(restart_in + 1) * ld;
    double* w = x + ld;
    double* r = w + ld;
    KrMat M{Kall + (size_t)b * ld * ld, ld, N, trans};

    const double bnrm2 = sqrt(kr_dot(rhs, rhs, N, red));
    const double atol = fmax(atol_in, rtol * bnrm2);
    int total = 0;
    if (bnrm2 == 0.0) {  // iterative.py: "if bnrm2 == 0: return b, 0"
        if (tid == 0) { info[b] = 0; if (iters) iters[b] = 0; }
        return;
    }
    // x = x0 (make_system copies it) or zeros
    const bool have_x0 = x0_all != nullptr || mask_all != nullptr;
    bool xany = false;
    {
        int any = 0;
        for (int i = tid; i < N; i += KR_THREADS) {
            double v = 0.0;
            if (x0_all != nullptr) v = x0_all[(size_t)b * ldr + i];
            else if (mask_all != nullptr && i < nmask && mask_all[(size_t)b * nmask + i]) v = rhs[i];
            x[i] = v;
            xs[i] = v;
            any |= (v != 0.0);
        }
        xany = __syncthreads_or(any) != 0;
    }
    double rnorm;
    if (xany) {  // r = b - A x, else b.copy()
        kr_matvec(M, xs, w);
        ++total;
        for (int i = tid; i < N; i += KR_THREADS) r[i] = rhs[i] - w[i];
    } else {
        for (int i = tid; i < N; i += KR_THREADS) r[i] = rhs[i];
    }
    if (have_x0) {  // gmres_solver.py:22-25 (the wrapper's workaround): |rhs - mat x0|_inf < atol returns x0
        double mx = 0.0;
        for (int i = tid; i < N; i += KR_THREADS) mx = fmax(mx, fabs(r[i]));
        mx = block_max(mx, red);
        if (mx < atol_in) {
            for (int i = tid; i < N; i += KR_THREADS) rhs[i] = x[i];
            if (tid == 0) { info[b] = 0; if (iters) iters[b] = 0; }  // the wrapper's own product is not scipy's
            return;
        }
    }
    rnorm = sqrt(kr_dot(r, r, N, red));
    if (rnorm < atol) {  // "Are we done?"
        for (int i = tid; i < N; i += KR_THREADS) rhs[i] = x[i];
        if (tid == 0) { info[b] = 0; if (iters) iters[b] = total; }
        return;
    }

    double ptol_max_factor = 1.0;
    double ptol = bnrm2 * fmin(ptol_max_factor, atol / bnrm2);
    double presid = 0.0;
    const int hs = restart_in + 1;  // row stride of hh

    for (int iteration = 0; iteration < maxiter; ++iteration) {
        // v[0] = r / |r|
        const double tmp0 = sqrt(kr_dot(r, r, N, red));
        {
            const double inv = 1.0 / tmp0;
            for (int i = tid; i < N; i += KR_THREADS) {
                const double v = r[i] * inv;
                V[i] = v;
                xs[i] = v;
            }
        }
        if (tid == 0) {
            for (int k = 0; k <= restart; ++k) S[k] = 0.0;
            S[0] = tmp0;
        }
        __syncthreads();
        bool breakdown = false;
        int col = 0;
        for (col = 0; col < restart; ++col) {
            // xs holds v[col]
            kr_matvec(M, xs, w);
            const double h0 = sqrt(kr_dot(w, w, N, red));
            for (int k = 0; k <= col; ++k) {  // modified Gram-Schmidt
                const double* vk = V + (size_t)k * ld;
                const double t = kr_dot(vk, w, N, red);
                if (tid == 0) hh[col * hs + k] = t;
                for (int i = tid; i < N; i += KR_THREADS) w[i] -= t * vk[i];
            }
            const double h1 = sqrt(kr_dot(w, w, N, red));
            double* vn = V + (size_t)(col + 1) * ld;
            double hcc1 = h1;
            if (h1 <= KR_EPS * h0) {  // exact solution indicator
                hcc1 = 0.0;
                breakdown = true;
                for (int i = tid; i < N; i += KR_THREADS) { vn[i] = w[i]; xs[i] = w[i]; }
            } else {
                const double inv = 1.0 / h1;
                for (int i = tid; i < N; i += KR_THREADS) {
                    const double v = w[i] * inv;
                    vn[i] = v;
                    xs[i] = v;
                }
            }
            if (tid == 0) {
                double* hc = hh + col * hs;
                hc[col + 1] = hcc1;
                for (int k = 0; k < col; ++k) {  // past rotations
                    const double c = giv[2 * k], s = giv[2 * k + 1];
                    const double n0 = hc[k], n1 = hc[k + 1];
                    hc[k] = c * n0 + s * n1;
                    hc[k + 1] = -s * n0 + c * n1;
                }
                double c, s, mag;
                kr_lartg(hc[col], hc[col + 1], c, s, mag);
                giv[2 * col] = c;
                giv[2 * col + 1] = s;
                hc[col] = mag;
                hc[col + 1] = 0.0;
                const double t = -s * S[col];
                S[col] = c * S[col];
                S[col + 1] = t;
                sc[0] = fabs(t);
            }
            __syncthreads();
            presid = sc[0];
            ++total;
            if (presid <= ptol || breakdown) break;
        }
        if (col == restart) col = restart - 1;  // Python's loop variable after a loop that ran to the end
        if (tid == 0) {
            if (hh[col * hs + col] == 0.0) S[col] = 0.0;
            for (int k = 0; k <= col; ++k) yv[k] = S[k];
            for (int k = col; k > 0; --k) {
                if (yv[k] != 0.0) {
                    yv[k] /= hh[k * hs + k];
                    const double t = yv[k];
                    for (int j = 0; j < k; ++j) yv[j] -= t * hh[k * hs + j];
                }
            }
            if (yv[0] != 0.0) yv[0] /= hh[0];
        }
        __syncthreads();
        for (int i = tid; i < N; i += KR_THREADS) {  // x += y @ v[:col+1]
            double acc = 0.0;
            for (int k = 0; k <= col; ++k) acc = fma(yv[k], V[(size_t)k * ld + i], acc);
            const double v = x[i] + acc;
            x[i] = v;
            xs[i] = v;
        }
        __syncthreads();
        kr_matvec(M, xs, w);
        ++total;
        for (int i = tid; i < N; i += KR_THREADS) r[i] = rhs[i] - w[i];
        rnorm = sqrt(kr_dot(r, r, N, red));
        if (rnorm <= atol) break;
        if (breakdown) break;
        if (presid <= ptol) ptol_max_factor = fmax(KR_EPS, 0.25 * ptol_max_factor);
        else ptol_max_factor = fmin(1.0, 1.5 * ptol_max_factor);
        ptol = presid * fmin(ptol_max_factor, atol / rnorm);
    }
    __syncthreads();
    for (int i = tid; i < N; i += KR_THREADS) rhs[i] = x[i];
    if (tid == 0) {
        info[b] = rnorm <= atol ? 0 : maxiter;  // gmres_solver.py:32-33: info != 0 raises LinearSolverError
        if (iters) iters[b] = total;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// MINRES (shift = 0, no preconditioner): scratch per instance = 7 rows of ld doubles (y, r1, r2, w, w1, w2, x).
__global__ void __launch_bounds__(KR_THREADS) minres_kernel(int ld, int ldr, int Nmax, const int32_t* __restrict__ Nvec,
                                                            const double* __restrict__ Kall, double* __restrict__ rhs_all,
                                                            const double* __restrict__ x0_all, double rtol,
                                                            double* __restrict__ scratch_all, int32_t* __restrict__ info,
                                                            int32_t* __restrict__ iters, GfWork work) {
    const int b = gf_instance(work, blockIdx.x);
    if (b < 0) return;
    const int N = Nvec != nullptr ? Nvec[b] : Nmax;
    const int tid = threadIdx.x;
    extern __shared__ double sm[];
    double* vs = sm;          // Nmax: the Lanczos vector v (input of the product, kept for the w update)
    double* red = vs + Nmax;  // 32
    double* rhs = rhs_all + (size_t)b * ldr;
    if (N <= 0) {
        if (tid == 0) { info[b] = 0; if (iters) iters[b] = 0; }
        return;
    }
    const int maxiter = 5 * N;
    double* base = scratch_all + (size_t)b * 7 * ld;
    double* y = base;
    double* r1 = base + ld;
    double* r2 = base + 2 * (size_t)ld;
    double* w = base + 3 * (size_t)ld;
    double* w1 = base + 4 * (size_t)ld;
    double* w2 = base + 5 * (size_t)ld;
    double* x = base + 6 * (size_t)ld;
    KrMat M{Kall + (size_t)b * ld * ld, ld, N, 0};

    // r1 = b - A x0 (or b), y = r1
    int nprod = 0;
    if (x0_all != nullptr) {
        for (int i = tid; i < N; i += KR_THREADS) {
            const double v = x0_all[(size_t)b * ldr + i];
            x[i] = v;
            vs[i] = v;
        }
        __syncthreads();
        kr_matvec(M, vs, y);
        nprod = 1;
        for (int i = tid; i < N; i += KR_THREADS) {
            const double v = rhs[i] - y[i];
            r1[i] = v;
            y[i] = v;
        }
    } else {
        for (int i = tid; i < N; i += KR_THREADS) {
            x[i] = 0.0;
            r1[i] = rhs[i];
            y[i] = rhs[i];
        }
    }
    double beta1 = kr_dot(r1, y, N, red);
    const double bnorm = sqrt(kr_dot(rhs, rhs, N, red));
    if (beta1 == 0.0 || bnorm == 0.0) {  // x0 is exact / b = 0: "return (x, 0)"
        for (int i = tid; i < N; i += KR_THREADS) rhs[i] = (beta1 == 0.0) ? x[i] : rhs[i];
        if (tid == 0) { info[b] = 0; if (iters) iters[b] = nprod; }
        return;
    }
    beta1 = sqrt(beta1);
    double oldb = 0.0, beta = beta1, dbar = 0.0, epsln = 0.0, phibar = beta1;
    double tnorm2 = 0.0, gmax = 0.0, gmin = 1.7976931348623157e308, cs = -1.0, sn = 0.0;
    for (int i = tid; i < N; i += KR_THREADS) { w[i] = 0.0; w2[i] = 0.0; w1[i] = 0.0; }
    // r2 = r1 (the same array in scipy until the first "r1 = r2; r2 = y")
    for (int i = tid; i < N; i += KR_THREADS) r2[i] = r1[i];
    int istop = 0, itn = 0;
    while (itn < maxiter) {
        ++itn;
        const double s = 1.0 / beta;
        for (int i = tid; i < N; i += KR_THREADS) vs[i] = s * y[i];
        __syncthreads();
        kr_matvec(M, vs, y);  // y = A v  (shift = 0)
        if (itn >= 2) {
            const double f = beta / oldb;
            for (int i = tid; i < N; i += KR_THREADS) y[i] = y[i] - f * r1[i];
        }
        const double alfa = kr_dot(vs, y, N, red);
        {
            const double f = alfa / beta;
            for (int i = tid; i < N; i += KR_THREADS) y[i] = y[i] - f * r2[i];
        }
        // r1 = r2; r2 = y; y = psolve(r2) = r2: rotate the three buffers (y becomes a copy of the new r2)
        {
            double* t = r1;
            r1 = r2;
            r2 = y;
            y = t;
            for (int i = tid; i < N; i += KR_THREADS) y[i] = r2[i];
        }
        oldb = beta;
        beta = kr_dot(r2, y, N, red);
        beta = sqrt(beta);
        tnorm2 += alfa * alfa + oldb * oldb + beta * beta;
        if (itn == 1 && beta / beta1 <= 10.0 * KR_EPS) istop = -1;
        const double oldeps = epsln;
        const double delta = cs * dbar + sn * alfa;
        const double gbar = sn * dbar - cs * alfa;
        epsln = sn * beta;
        dbar = -cs * beta;
        const double root = sqrt(gbar * gbar + dbar * dbar);
        double gamma = sqrt(gbar * gbar + beta * beta);
        gamma = fmax(gamma, KR_EPS);
        cs = gbar / gamma;
        sn = beta / gamma;
        const double phi = cs * phibar;
        phibar = sn * phibar;
        const double denom = 1.0 / gamma;
        {   // w1 = w2; w2 = w; w = (v - oldeps w1 - delta w2) denom; x += phi w
            double* t = w1;
            w1 = w2;
            w2 = w;
            w = t;
            for (int i = tid; i < N; i += KR_THREADS) {
                const double wn = (vs[i] - oldeps * w1[i] - delta * w2[i]) * denom;
                w[i] = wn;
                x[i] = x[i] + phi * wn;
            }
        }
        gmax = fmax(gmax, gamma);
        gmin = fmin(gmin, gamma);
        const double Anorm = sqrt(tnorm2);
        const double ynorm = sqrt(kr_dot(x, x, N, red));
        const double epsx = Anorm * ynorm * KR_EPS;
        const double rnorm = phibar;
        const double test1 = (ynorm == 0.0 || Anorm == 0.0) ? INFINITY : rnorm / (Anorm * ynorm);
        const double test2 = (Anorm == 0.0) ? INFINITY : root / Anorm;
        const double Acond = gmax / gmin;
        if (istop == 0) {
            const double t1 = 1.0 + test1, t2 = 1.0 + test2;
            if (t2 <= 1.0) istop = 2;
            if (t1 <= 1.0) istop = 1;
            if (itn >= maxiter) istop = 6;
            if (Acond >= 0.1 / KR_EPS) istop = 4;
            if (epsx >= beta1) istop = 3;
            if (test2 <= rtol) istop = 2;
            if (test1 <= rtol) istop = 1;
        }
        if (istop != 0) break;
    }
    __syncthreads();
    for (int i = tid; i < N; i += KR_THREADS) rhs[i] = x[i];
    if (tid == 0) {
        info[b] = istop == 6 ? maxiter : 0;  // minres_solver.py:21-22
        if (iters) iters[b] = itn + nprod;
    }
}

}  // namespace

extern "C" {

int gf_krylov_scratch_rows(int method, int restart) {
    return method == GF_KRYLOV_MINRES ? 7 : (restart > 0 ? restart : 20) + 4;
}

int gf_gmres_solve(int B, int ld, int Nmax, const int32_t* Nvec, const double* K, double* rhs, int ldr, const double* x0,
                   const uint8_t* x0_mask, int nmask, int trans, int restart, double rtol, double atol, double* scratch,
                   int32_t* info, int32_t* iters, const int32_t* work, const int32_t* nwork_dev, int nwork,
                   void* stream) {
    if (B < 0 || ld < 1 || Nmax < 0 || Nmax > ld || ldr < Nmax || nwork < 0 || restart < 1 || restart > 64)
        return GF_ERR_ARG;
    if (nwork == 0 || B == 0) return GF_OK;
    const size_t smem = sizeof(double) * ((size_t)Nmax + 32 + (size_t)restart * (restart + 1) + 2 * restart +
                                          2 * (restart + 1) + 4);
    if (smem > 200 * 1024) return GF_ERR_UNSUPPORTED;
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(gmres_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        attr = true;
    }
    gmres_kernel<<<nwork, KR_THREADS, smem, (cudaStream_t)stream>>>(ld, ldr, Nmax, Nvec, K, rhs, x0, x0_mask, nmask, trans,
                                                                    restart, rtol, atol, scratch, info, iters,
                                                                    GfWork{work, nwork_dev});
    return gf_launch_status();
}

int gf_minres_solve(int B, int ld, int Nmax, const int32_t* Nvec, const double* K, double* rhs, int ldr, const double* x0,
                    double rtol, double* scratch, int32_t* info, int32_t* iters, const int32_t* work,
                    const int32_t* nwork_dev, int nwork, void* stream) {
    if (B < 0 || ld < 1 || Nmax < 0 || Nmax > ld || ldr < Nmax || nwork < 0) return GF_ERR_ARG;
    if (nwork == 0 || B == 0) return GF_OK;
    const size_t smem = sizeof(double) * ((size_t)Nmax + 32);
    if (smem > 200 * 1024) return GF_ERR_UNSUPPORTED;
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(minres_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        attr = true;
    }
    minres_kernel<<<nwork, KR_THREADS, smem, (cudaStream_t)stream>>>(ld, ldr, Nmax, Nvec, K, rhs, x0, rtol, scratch, info,
                                                                     iters, GfWork{work, nwork_dev});
    return gf_launch_status();
}

}  // extern "C"
