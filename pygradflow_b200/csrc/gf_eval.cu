// Problem-family evaluators and the augmented-Lagrangian gradient (a17).
//
// Reference: pygradflow/iterate.py:59-110 (cached obj_grad / cons / cons_jac, aug_lag_deriv_x),
// tests/pygradflow/qp.py:17-30 (QP callbacks), tests/pygradflow/rosenbrock.py:15-46 (n = 2 Rosenbrock).
//
// All kernels: one CTA per instance, HBM-bound streaming over H / A / J with coalesced column-per-thread
// accumulation in ascending row order with un-fused multiply/add -- the same summation order scipy's
// csc_matvec uses for J'v and Hx, so p = lamb*x^ - grad L (which decides the active set) is reproduced
// to the last bit wherever the CPU side is sequential too.
#include "gf_common.cuh"
#include "../../include/gradflow_b200.h"

namespace {

// grad = H x + g (H symmetric: column i is read as row i -> coalesced), cons = A x + b,
// obj = 1/2 x'Hx + g'x.
__global__ void qp_eval_kernel(int n, int m, const double* __restrict__ H, const double* __restrict__ A,
                               const double* __restrict__ g, const double* __restrict__ bvec,
                               const double* __restrict__ x, double* __restrict__ grad,
                               double* __restrict__ cons, double* __restrict__ obj, GfWork work) {
    const int b = gf_instance(work, blockIdx.x);
    if (b < 0) return;
    extern __shared__ double sm[];
    double* xs = sm;           // n
    double* red = sm + n;      // 32
    const double* Hb = H + (size_t)b * n * n;
    const double* xb = x + (size_t)b * n;
    for (int i = threadIdx.x; i < n; i += blockDim.x) xs[i] = xb[i];
    __syncthreads();
    double objp = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double acc = 0.0;
        const double* col = Hb + i;
        int j = 0;
        for (; j + 8 <= n; j += 8) {
            double h[8];
#pragma unroll
            for (int u = 0; u < 8; u++) h[u] = __ldg(col + (size_t)(j + u) * n);
#pragma unroll
            for (int u = 0; u < 8; u++) acc = __dadd_rn(acc, __dmul_rn(h[u], xs[j + u]));
        }
        for (; j < n; j++) acc = __dadd_rn(acc, __dmul_rn(__ldg(col + (size_t)j * n), xs[j]));
        const double gi = g[(size_t)b * n + i];
        grad[(size_t)b * n + i] = acc + gi;
        objp += xs[i] * (0.5 * acc + gi);
    }
    if (obj != nullptr) {
        const double o = block_sum(objp, red);
        if (threadIdx.x == 0) obj[b] = o;
    }
    if (m > 0) {
        const double* Ab = A + (size_t)b * m * n;
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
        for (int r = wid; r < m; r += nw) {
            const double* row = Ab + (size_t)r * n;
            double acc = 0.0;
            for (int j = lane; j < n; j += 32) acc += __ldg(row + j) * xs[j];
            acc = warp_sum(acc);
            if (lane == 0) cons[(size_t)b * m + r] = acc + bvec[(size_t)b * m + r];
        }
    }
}

// dL = grad + J'(rho c + y)   (iterate.py:91-94);  jty = J'y (iterate.py:138,171);  jtc = J'c (iterate.py:125)
__global__ void aug_lag_grad_kernel(int n, int m, const double* __restrict__ J, const double* __restrict__ grad,
                                    const double* __restrict__ cons, const double* __restrict__ y,
                                    const double* __restrict__ rho, double* __restrict__ dL,
                                    double* __restrict__ jty, double* __restrict__ jtc, GfWork work) {
    const int b = gf_instance(work, blockIdx.x);
    if (b < 0) return;
    extern __shared__ double sm[];
    double* v = sm;          // rho c + y
    double* ys = sm + m;
    double* cs = sm + 2 * m;
    const double r = rho[b];
    for (int j = threadIdx.x; j < m; j += blockDim.x) {
        const double c = cons[(size_t)b * m + j], yy = y[(size_t)b * m + j];
        v[j] = __dadd_rn(__dmul_rn(r, c), yy);
        ys[j] = yy;
        cs[j] = c;
    }
    __syncthreads();
    const double* Jb = J + (size_t)b * m * n;
    const bool want_y = jty != nullptr, want_c = jtc != nullptr;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double a1 = 0.0, a2 = 0.0, a3 = 0.0;
        const double* col = Jb + i;
        int j = 0;
        for (; j + 8 <= m; j += 8) {
            double h[8];
#pragma unroll
            for (int u = 0; u < 8; u++) h[u] = __ldg(col + (size_t)(j + u) * n);
#pragma unroll
            for (int u = 0; u < 8; u++) {
                a1 = __dadd_rn(a1, __dmul_rn(h[u], v[j + u]));
                if (want_y) a2 = __dadd_rn(a2, __dmul_rn(h[u], ys[j + u]));
                if (want_c) a3 = __dadd_rn(a3, __dmul_rn(h[u], cs[j + u]));
            }
        }
        for (; j < m; j++) {
            const double h = __ldg(col + (size_t)j * n);
            a1 = __dadd_rn(a1, __dmul_rn(h, v[j]));
            if (want_y) a2 = __dadd_rn(a2, __dmul_rn(h, ys[j]));
            if (want_c) a3 = __dadd_rn(a3, __dmul_rn(h, cs[j]));
        }
        const size_t o = (size_t)b * n + i;
        if (dL != nullptr) dL[o] = __dadd_rn(grad[o], a1);
        if (want_y) jty[o] = a2;
        if (want_c) jtc[o] = a3;
    }
}

// Chained Rosenbrock f = sum_i b_i (x_{i+1} - x_i^2)^2 + (a_i - x_i)^2, i = 0..n-2.
// Operation order follows the NumPy expressions of the oracle / golden problem class term by term.
__global__ void rosen_eval_kernel(int n, const double* __restrict__ a, const double* __restrict__ bc,
                                  const double* __restrict__ x, double* __restrict__ grad,
                                  double* __restrict__ obj, GfWork work) {
    const int b = gf_instance(work, blockIdx.x);
    if (b < 0) return;
    __shared__ double red[32];
    const double* xb = x + (size_t)b * n;
    const double* ab = a + (size_t)b * (n - 1);
    const double* bb = bc + (size_t)b * (n - 1);
    double objp = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double xi = xb[i];
        double gi = 0.0;
        if (i < n - 1) {
            const double r = __dsub_rn(xb[i + 1], __dmul_rn(xi, xi));
            const double d = __dsub_rn(ab[i], xi);
            const double t1 = __dmul_rn(__dmul_rn(__dmul_rn(-4.0, bb[i]), r), xi);
            gi = __dadd_rn(gi, __dsub_rn(t1, __dmul_rn(2.0, d)));
            objp += __dadd_rn(__dmul_rn(__dmul_rn(bb[i], r), r), __dmul_rn(d, d));
        }
        if (i > 0) {
            const double xm = xb[i - 1];
            const double rm = __dsub_rn(xi, __dmul_rn(xm, xm));
            gi = __dadd_rn(gi, __dmul_rn(__dmul_rn(2.0, bb[i - 1]), rm));
        }
        grad[(size_t)b * n + i] = gi;
    }
    if (obj != nullptr) {
        const double o = block_sum(objp, red);
        if (threadIdx.x == 0) obj[b] = o;
    }
}

// Writes the three diagonals of the dense Hessian H[b] (n x n, row-major); all other entries must
// already be zero (the caller zero-fills H once).
__global__ void rosen_hess_kernel(int n, const double* __restrict__ a, const double* __restrict__ bc,
                                  const double* __restrict__ x, double* __restrict__ H, GfWork work) {
    const int b = gf_instance(work, blockIdx.x);
    if (b < 0) return;
    const double* xb = x + (size_t)b * n;
    const double* bb = bc + (size_t)b * (n - 1);
    double* Hb = H + (size_t)b * n * n;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double xi = xb[i];
        double main = 0.0;
        if (i < n - 1) {
            const double r = __dsub_rn(xb[i + 1], __dmul_rn(xi, xi));
            const double t = __dsub_rn(__dmul_rn(__dmul_rn(8.0, bb[i]), __dmul_rn(xi, xi)),
                                       __dmul_rn(__dmul_rn(4.0, bb[i]), r));
            main = __dadd_rn(main, __dadd_rn(t, 2.0));
            const double off = __dmul_rn(__dmul_rn(-4.0, bb[i]), xi);
            Hb[(size_t)i * n + i + 1] = off;
            Hb[(size_t)(i + 1) * n + i] = off;
        }
        if (i > 0) main = __dadd_rn(main, __dmul_rn(2.0, bb[i - 1]));
        Hb[(size_t)i * n + i] = main;
    }
}

// ---------------------------------------------------------------------------------------------
// OCP family (cfg4): z = (x_1, u_0, ..., x_S, u_{S-1}), dynamics c_j = x_{j+1} - x_j - h (A_j x_j + B_j u_j +
// 0.1 sin x_j), cost 1/2 sum (x'Qx + u'Ru).  Same accumulation order as the oracle's OCP class (sequential in k).
__global__ void ocp_eval_kernel(int S, int nx, int nu, double h, const double* __restrict__ A,
                                const double* __restrict__ Bm, const double* __restrict__ Q,
                                const double* __restrict__ R, const double* __restrict__ xinit,
                                const double* __restrict__ z, double* __restrict__ grad, double* __restrict__ cons,
                                double* __restrict__ obj, GfWork work) {
    const int b = gf_instance(work, blockIdx.x);
    if (b < 0) return;
    __shared__ double red[32];
    const int w = nx + nu, n = S * w, m = S * nx;
    const double* zb = z + (size_t)b * n;
    const double* Ab = A + (size_t)b * S * nx * nx;
    const double* Bb = Bm + (size_t)b * S * nx * nu;
    const double* Qb = Q + (size_t)b * S * nx;
    const double* Rb = R + (size_t)b * S * nu;
    const double* x0 = xinit + (size_t)b * nx;
    double objp = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int j = i / w, c = i - j * w;
        const double v = zb[i];
        const double wgt = c < nx ? Qb[j * nx + c] : Rb[j * nu + (c - nx)];
        const double gi = __dmul_rn(wgt, v);
        grad[(size_t)b * n + i] = gi;
        objp += 0.5 * gi * v;
    }
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
        const int j = i / nx, r = i - j * nx;
        const double* xp = j == 0 ? x0 : zb + (size_t)(j - 1) * w;  // x_j
        const double* uj = zb + (size_t)j * w + nx;
        const double* Ar = Ab + ((size_t)j * nx + r) * nx;
        const double* Br = Bb + ((size_t)j * nx + r) * nu;
        double ax = 0.0, bu = 0.0;
        for (int k = 0; k < nx; k++) ax = __dadd_rn(ax, __dmul_rn(Ar[k], xp[k]));
        for (int k = 0; k < nu; k++) bu = __dadd_rn(bu, __dmul_rn(Br[k], uj[k]));
        const double f = __dadd_rn(__dadd_rn(ax, bu), __dmul_rn(0.1, sin(xp[r])));
        cons[(size_t)b * m + i] = __dsub_rn(__dsub_rn(zb[(size_t)j * w + r], xp[r]), __dmul_rn(h, f));
    }
    if (obj != nullptr) {
        const double o = block_sum(objp, red);
        if (threadIdx.x == 0) obj[b] = o;
    }
}

// The non-zero blocks of the dense Jacobian J[b] (m x n, zero elsewhere): d c_j / d x_{j+1} = I,
// d c_j / d u_j = -h B_j, d c_j / d x_j = -(I + h (A_j + 0.1 diag cos x_j)).
__global__ void ocp_jac_kernel(int S, int nx, int nu, double h, const double* __restrict__ A,
                               const double* __restrict__ Bm, const double* __restrict__ z,
                               double* __restrict__ J, GfWork work) {
    const int b = gf_instance(work, blockIdx.x);
    if (b < 0) return;
    const int w = nx + nu, n = S * w, m = S * nx;
    const double* zb = z + (size_t)b * n;
    const double* Ab = A + (size_t)b * S * nx * nx;
    const double* Bb = Bm + (size_t)b * S * nx * nu;
    double* Jb = J + (size_t)b * m * n;
    const int per = nx * (2 * nx + nu);  // entries per stage: [x_j | x_{j+1} | u_j] columns of nx rows
    for (int e = threadIdx.x; e < S * per; e += blockDim.x) {
        const int j = e / per, t = e - j * per, r = t / (2 * nx + nu), c = t - r * (2 * nx + nu);
        double* row = Jb + (size_t)(j * nx + r) * n;
        if (c < nx) {  // d / d x_j (previous stage block), j >= 1
            if (j >= 1) {
                const double d = (r == c) ? 1.0 : 0.0;
                const double xe = zb[(size_t)(j - 1) * w + c];
                const double ee = __dmul_rn(h, __dadd_rn(Ab[((size_t)j * nx + r) * nx + c],
                                                        __dmul_rn(d, __dmul_rn(0.1, cos(xe)))));
                row[(size_t)(j - 1) * w + c] = -__dadd_rn(d, ee);
            }
        } else if (c < 2 * nx) {
            row[(size_t)j * w + (c - nx)] = (r == c - nx) ? 1.0 : 0.0;
        } else {
            row[(size_t)j * w + nx + (c - 2 * nx)] = -__dmul_rn(h, Bb[((size_t)j * nx + r) * nu + (c - 2 * nx)]);
        }
    }
}

// Diagonal of the Hessian of the Lagrangian (H must be zero elsewhere):
// x_{j+1}: Q_j + (y_{j+1} * 0.1 h) sin(x_{j+1}) (no dynamics term for x_S), u_j: R_j.
__global__ void ocp_hess_kernel(int S, int nx, int nu, double c1, const double* __restrict__ Q,
                                const double* __restrict__ R, const double* __restrict__ z,
                                const double* __restrict__ y, double* __restrict__ H, GfWork work) {
    const int b = gf_instance(work, blockIdx.x);
    if (b < 0) return;
    const int w = nx + nu, n = S * w, m = S * nx;
    const double* zb = z + (size_t)b * n;
    const double* yb = y + (size_t)b * m;
    double* Hb = H + (size_t)b * n * n;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int j = i / w, c = i - j * w;
        double v;
        if (c < nx) {
            v = Q[((size_t)b * S + j) * nx + c];
            if (j + 1 < S) v = __dadd_rn(v, __dmul_rn(__dmul_rn(yb[(size_t)(j + 1) * nx + c], c1), sin(zb[i])));
        } else {
            v = R[((size_t)b * S + j) * nu + (c - nx)];
        }
        Hb[(size_t)i * n + i] = v;
    }
}

inline int pick_threads(int n) {
    int t = ((n + 31) / 32) * 32;
    if (t < 64) t = 64;
    if (t > 512) t = 512;
    return t;
}

}  // namespace

extern "C" int gf_qp_eval(int B, int n, int m, const double* H, const double* A, const double* g, const double* b,
                          const double* x, double* grad, double* cons, double* obj, const int32_t* work,
                          const int32_t* nwork_dev, int nwork, void* stream) {
    if (B <= 0 || n <= 0 || m < 0 || !H || !g || !x || !grad || (m > 0 && (!A || !b || !cons))) return GF_ERR_ARG;
    if (nwork <= 0) return GF_OK;
    const size_t smem = (size_t)(n + 32) * sizeof(double);
    if (smem > 200 * 1024) return GF_ERR_UNSUPPORTED;
    if (smem > 48 * 1024) cudaFuncSetAttribute(qp_eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    qp_eval_kernel<<<nwork, pick_threads(n), smem, (cudaStream_t)stream>>>(n, m, H, A, g, b, x, grad, cons, obj,
                                                                            GfWork{work, nwork_dev});
    return gf_launch_status();
}

extern "C" int gf_aug_lag_grad(int B, int n, int m, const double* J, const double* grad, const double* cons,
                               const double* y, const double* rho, double* dL, double* jty, double* jtc,
                               const int32_t* work, const int32_t* nwork_dev, int nwork, void* stream) {
    if (B <= 0 || n <= 0 || m < 0 || !grad || !rho || (m > 0 && (!J || !cons || !y))) return GF_ERR_ARG;
    if (nwork <= 0) return GF_OK;
    const size_t smem = (size_t)(3 * m + 1) * sizeof(double);
    if (smem > 200 * 1024) return GF_ERR_UNSUPPORTED;
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(aug_lag_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    aug_lag_grad_kernel<<<nwork, pick_threads(n), smem, (cudaStream_t)stream>>>(n, m, J, grad, cons, y, rho, dL, jty,
                                                                                 jtc, GfWork{work, nwork_dev});
    return gf_launch_status();
}

extern "C" int gf_rosen_eval(int B, int n, const double* a, const double* b, const double* x, double* grad,
                             double* obj, const int32_t* work, const int32_t* nwork_dev, int nwork, void* stream) {
    if (B <= 0 || n < 2 || !a || !b || !x || !grad) return GF_ERR_ARG;
    if (nwork <= 0) return GF_OK;
    rosen_eval_kernel<<<nwork, pick_threads(n), 0, (cudaStream_t)stream>>>(n, a, b, x, grad, obj,
                                                                            GfWork{work, nwork_dev});
    return gf_launch_status();
}

extern "C" int gf_rosen_hess(int B, int n, const double* a, const double* b, const double* x, double* H,
                             const int32_t* work, const int32_t* nwork_dev, int nwork, void* stream) {
    if (B <= 0 || n < 2 || !a || !b || !x || !H) return GF_ERR_ARG;
    if (nwork <= 0) return GF_OK;
    rosen_hess_kernel<<<nwork, pick_threads(n), 0, (cudaStream_t)stream>>>(n, a, b, x, H, GfWork{work, nwork_dev});
    return gf_launch_status();
}

extern "C" int gf_ocp_eval(int B, int S, int nx, int nu, double h, const double* A, const double* Bm, const double* Q,
                           const double* R, const double* xinit, const double* z, double* grad, double* cons,
                           double* obj, const int32_t* work, const int32_t* nwork_dev, int nwork, void* stream) {
    if (B <= 0 || S <= 0 || nx <= 0 || nu <= 0 || !A || !Bm || !Q || !R || !xinit || !z || !grad || !cons)
        return GF_ERR_ARG;
    if (nwork <= 0) return GF_OK;
    ocp_eval_kernel<<<nwork, pick_threads(S * (nx + nu)), 0, (cudaStream_t)stream>>>(
        S, nx, nu, h, A, Bm, Q, R, xinit, z, grad, cons, obj, GfWork{work, nwork_dev});
    return gf_launch_status();
}

extern "C" int gf_ocp_jac(int B, int S, int nx, int nu, double h, const double* A, const double* Bm, const double* z,
                          double* J, const int32_t* work, const int32_t* nwork_dev, int nwork, void* stream) {
    if (B <= 0 || S <= 0 || nx <= 0 || nu <= 0 || !A || !Bm || !z || !J) return GF_ERR_ARG;
    if (nwork <= 0) return GF_OK;
    ocp_jac_kernel<<<nwork, 512, 0, (cudaStream_t)stream>>>(S, nx, nu, h, A, Bm, z, J, GfWork{work, nwork_dev});
    return gf_launch_status();
}

extern "C" int gf_ocp_hess(int B, int S, int nx, int nu, double c1, const double* Q, const double* R, const double* z,
                           const double* y, double* H, const int32_t* work, const int32_t* nwork_dev, int nwork,
                           void* stream) {
    if (B <= 0 || S <= 0 || nx <= 0 || nu <= 0 || !Q || !R || !z || !y || !H) return GF_ERR_ARG;
    if (nwork <= 0) return GF_OK;
    ocp_hess_kernel<<<nwork, pick_threads(S * (nx + nu)), 0, (cudaStream_t)stream>>>(S, nx, nu, c1, Q, R, z, y, H,
                                                                                       GfWork{work, nwork_dev});
    return gf_launch_status();
}
