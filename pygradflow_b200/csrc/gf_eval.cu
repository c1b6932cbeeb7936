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
