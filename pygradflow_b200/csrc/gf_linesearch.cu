// Globalized Newton: merit gradient and the fused residual-norm / Armijo reduction (a3, K4).
//
// Reference: pygradflow/newton.py:242-304 (GlobalizedNewtonMethod.step).  The reference assembles the full
// (n+m)^2 matrix F' (implicit_func.py:254-294) only to form F'^T F (newton.py:262-263, with a TODO asking for a
// forward product); here the product is computed matrix-free:
//     v = P_I F_x (active rows zeroed, implicit_func.py:80-99),  u = J v,
//     grad_x = lamb F_x + H_rho v - J' F_y,   H_rho v = H v + rho J'(J v)   (iterate.py:102-110),
//     grad_y = u + lamb F_y,
// and only the two scalars the line search needs leave the kernel: res = 1/2 |F|^2 and
// inner = grad . (dx, dy) (newton.py:254,271).
#include "gf_common.cuh"
#include "../../include/gradflow_b200.h"

namespace {

__global__ void merit_grad_kernel(int n, int m, const double* __restrict__ H, const double* __restrict__ J,
                                  const double* __restrict__ F, const uint8_t* __restrict__ active,
                                  const double* __restrict__ dt, const double* __restrict__ rho,
                                  const double* __restrict__ dx, const double* __restrict__ dy,
                                  double* __restrict__ res, double* __restrict__ inner, GfWork work) {
    const int b = gf_instance(work, blockIdx.x);
    if (b < 0) return;
    extern __shared__ double sm[];
    double* v = sm;            // n : P_I F_x
    double* z = sm + n;        // m : rho u - F_y
    double* red = z + m;       // 32
    const double lamb = 1.0 / dt[b];
    const double rh = rho[b];
    const double* Fb = F + (size_t)b * (n + m);
    double ss = 0.0, ip = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double fx = Fb[i];
        v[i] = active[(size_t)b * n + i] ? 0.0 : fx;
        ss += fx * fx;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const double* Jb = J + (size_t)b * m * n;
    for (int j = wid; j < m; j += nw) {
        const double* row = Jb + (size_t)j * n;
        double acc = 0.0;
        int i = lane;
        for (; i + 96 < n; i += 128) {  // four loads in flight per lane
            const double r0 = __ldg(row + i), r1 = __ldg(row + i + 32), r2 = __ldg(row + i + 64), r3 = __ldg(row + i + 96);
            acc += r0 * v[i] + r1 * v[i + 32] + r2 * v[i + 64] + r3 * v[i + 96];
        }
        for (; i < n; i += 32) acc += __ldg(row + i) * v[i];
        acc = warp_sum(acc);  // u_j = (J v)_j
        if (lane == 0) {
            const double fy = Fb[n + j];
            z[j] = rh * acc - fy;
            ss += fy * fy;
            ip += (acc + lamb * fy) * dy[(size_t)b * m + j];  // grad_y . dy
        }
    }
    __syncthreads();
    const double* Hb = H + (size_t)b * n * n;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double acc = lamb * Fb[i];
        const double* hc = Hb + i;  // H symmetric: column i read as row i
        int j = 0;
        for (; j + 8 <= n; j += 8) {  // eight loads in flight per thread
            double h[8];
#pragma unroll
            for (int u = 0; u < 8; u++) h[u] = __ldg(hc + (size_t)(j + u) * n);
#pragma unroll
            for (int u = 0; u < 8; u++) acc = fma(h[u], v[j + u], acc);
        }
        for (; j < n; j++) acc = fma(__ldg(hc + (size_t)j * n), v[j], acc);
        const double* jc = Jb + i;
        j = 0;
        for (; j + 8 <= m; j += 8) {
            double h[8];
#pragma unroll
            for (int u = 0; u < 8; u++) h[u] = __ldg(jc + (size_t)(j + u) * n);
#pragma unroll
            for (int u = 0; u < 8; u++) acc = fma(h[u], z[j + u], acc);
        }
        for (; j < m; j++) acc = fma(__ldg(jc + (size_t)j * n), z[j], acc);
        ip += acc * dx[(size_t)b * n + i];
    }
    const double s2 = block_sum(ss, red);
    const double ipt = block_sum(ip, red);
    if (threadIdx.x == 0) {
        res[b] = 0.5 * s2;
        inner[b] = ipt;
    }
}

// trial point of the line search: xt = x - alpha dx, yt = y - alpha dy (un-clipped, newton.py:276-278);
// alpha = 1 uses (dx, dy) unscaled exactly like the reference's first trial.
__global__ void ls_trial_kernel(int n, int m, const double* __restrict__ x, const double* __restrict__ y,
                                const double* __restrict__ dx, const double* __restrict__ dy,
                                const double* __restrict__ alpha, double* __restrict__ xt, double* __restrict__ yt,
                                GfWork work) {
    const int b = gf_instance(work, blockIdx.x);
    if (b < 0) return;
    const double a = alpha[b];
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const size_t o = (size_t)b * n + i;
        xt[o] = __dsub_rn(x[o], a == 1.0 ? dx[o] : __dmul_rn(a, dx[o]));
    }
    for (int j = threadIdx.x; j < m; j += blockDim.x) {
        const size_t o = (size_t)b * m + j;
        yt[o] = __dsub_rn(y[o], a == 1.0 ? dy[o] : __dmul_rn(a, dy[o]));
    }
}

// Fused: scaled residual F at the trial point (active set recomputed, implicit_func.py:219-231), 1/2 |F|^2 by a
// warp-shuffle block reduction, and the Armijo decision of newton.py:280-290 applied in the same kernel:
//   state 0 = searching, 1 = accepted at the current alpha, 2 = exhausted (the reference raises, newton.py:294).
__global__ void armijo_residual_kernel(int n, int m, const double* __restrict__ xt, const double* __restrict__ yt,
                                       const double* __restrict__ x0, const double* __restrict__ y0,
                                       const double* __restrict__ dL, const double* __restrict__ cons,
                                       const double* __restrict__ lb, const double* __restrict__ ub,
                                       const double* __restrict__ dt, const double* __restrict__ res,
                                       const double* __restrict__ inner, double newton_tol, int max_trials,
                                       double* __restrict__ alpha, int32_t* __restrict__ trials,
                                       int32_t* __restrict__ state, double* __restrict__ next_res, GfWork work) {
    const int b = gf_instance(work, blockIdx.x);
    if (b < 0) return;
    if (state[b] != 0) return;
    __shared__ double red[32];
    const double lamb = 1.0 / dt[b];
    double ss = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const size_t o = (size_t)b * n + i;
        const double p = __dsub_rn(__dmul_rn(lamb, x0[o]), dL[o]);
        const double lo = __dmul_rn(lamb, lb[o]), hi = __dmul_rn(lamb, ub[o]);
        const bool act = (p < lo - GF_ACTIVE_SLACK) || (p > hi + GF_ACTIVE_SLACK);
        const double proj = act ? fmin(fmax(p, lo), hi) : p;
        const double rx = __dsub_rn(__dmul_rn(lamb, xt[o]), proj);
        ss += rx * rx;
    }
    for (int j = threadIdx.x; j < m; j += blockDim.x) {
        const size_t o = (size_t)b * m + j;
        const double ry = -__dsub_rn(__dmul_rn(lamb, yt[o]), __dadd_rn(__dmul_rn(lamb, y0[o]), cons[o]));
        ss += ry * ry;
    }
    const double tot = block_sum(ss, red);
    if (threadIdx.x == 0) {
        const double nres = 0.5 * tot;
        const double a = alpha[b];
        if (next_res != nullptr) next_res[b] = nres;
        const int t = trials[b] + 1;
        trials[b] = t;
        if (nres <= newton_tol || nres <= res[b] + (1e-4 * a * inner[b])) {  // note the '+' 
            state[b] = 1;
        } else if (t >= max_trials) {
            state[b] = 2;
        } else {
            alpha[b] = a * 0.5;
        }
    }
}

inline int pick_threads(int n) {
    int t = ((n + 31) / 32) * 32;
    if (t < 64) t = 64;
    if (t > 512) t = 512;
    return t;
}

}  // namespace

extern "C" int gf_merit_grad(int B, int n, int m, const double* H, const double* J, const double* F,
                             const uint8_t* active, const double* dt, const double* rho, const double* dx,
                             const double* dy, double* res, double* inner, const int32_t* work,
                             const int32_t* nwork_dev, int nwork, void* stream) {
    if (B <= 0 || n <= 0 || m < 0 || !H || !F || !active || !dt || !rho || !dx || !res || !inner) return GF_ERR_ARG;
    if (m > 0 && (!J || !dy)) return GF_ERR_ARG;
    if (nwork <= 0) return GF_OK;
    const size_t smem = (size_t)(n + m + 32) * sizeof(double);
    if (smem > 200 * 1024) return GF_ERR_UNSUPPORTED;
    if (smem > 48 * 1024) cudaFuncSetAttribute(merit_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    merit_grad_kernel<<<nwork, pick_threads(n), smem, (cudaStream_t)stream>>>(n, m, H, J, F, active, dt, rho, dx, dy,
                                                                              res, inner, GfWork{work, nwork_dev});
    return gf_launch_status();
}

// Start of the Armijo search (newton.py:256-257,273): instances of the work list whose merit value is already below
// newton_tol return the full step (state 1), the others search (state 0) from alpha = 1; instances outside the list are
// parked (state 3, set by the caller's fill before this launch).
__global__ void ls_begin_kernel(const double* __restrict__ res, double newton_tol, int32_t* __restrict__ state,
                                double* __restrict__ alpha, int32_t* __restrict__ trials, GfWork work, int nwork) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= nwork) return;
    const int b = gf_instance(work, w);
    if (b < 0) return;
    state[b] = res[b] <= newton_tol ? 1 : 0;
    alpha[b] = 1.0;
    trials[b] = 0;
}

extern "C" int gf_ls_begin(int B, const double* res, double newton_tol, int32_t* state, double* alpha, int32_t* trials,
                           const int32_t* work, const int32_t* nwork_dev, int nwork, void* stream) {
    if (B <= 0 || !res || !state || !alpha || !trials) return GF_ERR_ARG;
    if (nwork <= 0) return GF_OK;
    ls_begin_kernel<<<(nwork + 127) / 128, 128, 0, (cudaStream_t)stream>>>(res, newton_tol, state, alpha, trials,
                                                                          GfWork{work, nwork_dev}, nwork);
    return gf_launch_status();
}

extern "C" int gf_ls_trial(int B, int n, int m, const double* x, const double* y, const double* dx, const double* dy,
                           const double* alpha, double* xt, double* yt, const int32_t* work,
                           const int32_t* nwork_dev, int nwork, void* stream) {
    if (B <= 0 || n <= 0 || m < 0 || !x || !dx || !alpha || !xt) return GF_ERR_ARG;
    if (m > 0 && (!y || !dy || !yt)) return GF_ERR_ARG;
    if (nwork <= 0) return GF_OK;
    ls_trial_kernel<<<nwork, pick_threads(n), 0, (cudaStream_t)stream>>>(n, m, x, y, dx, dy, alpha, xt, yt,
                                                                         GfWork{work, nwork_dev});
    return gf_launch_status();
}

extern "C" int gf_armijo_residual(int B, int n, int m, const double* xt, const double* yt, const double* x0,
                                  const double* y0, const double* dL, const double* cons, const double* lb,
                                  const double* ub, const double* dt, const double* res, const double* inner,
                                  double newton_tol, int max_trials, double* alpha, int32_t* trials, int32_t* state,
                                  double* next_res, const int32_t* work, const int32_t* nwork_dev, int nwork,
                                  void* stream) {
    if (B <= 0 || n <= 0 || m < 0 || !xt || !x0 || !dL || !lb || !ub || !dt || !res || !inner || !alpha || !trials ||
        !state)
        return GF_ERR_ARG;
    if (m > 0 && (!yt || !y0 || !cons)) return GF_ERR_ARG;
    if (nwork <= 0) return GF_OK;
    armijo_residual_kernel<<<nwork, pick_threads(n + m), 0, (cudaStream_t)stream>>>(
        n, m, xt, yt, x0, y0, dL, cons, lb, ub, dt, res, inner, newton_tol, max_trials, alpha, trials, state, next_res,
        GfWork{work, nwork_dev});
    return gf_launch_status();
}
