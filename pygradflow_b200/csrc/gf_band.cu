// Banded LDL' for KKT systems whose coupling is local under a family-given ordering (cfg4: discretised optimal
// control, stage-interleaved primal / dual ordering, half-bandwidth nx + (nx + nu) - 1).
//
// Replaces, for such families, pygradflow/linear_solver/lu_solver.py:9-21 on the matrix of
// symmetric_step_solver.py:49-77.  The reference relies on SuperLU's sparse ordering for the same effect.
//
// Formulation: the FULL (n + m) KKT system in the family's order, with the rows / columns of active variables
// replaced by identity (their coupling is already on the right-hand side: symmetric_step_solver.py:79-94), so the
// order -- and the band -- never changes with the active set.  K is quasi-definite (H_II + lamb I > 0, -delta I < 0),
// hence L D L' exists without pivoting for ANY symmetric permutation and is stable; the signs of D give the inertia.
//
// Storage: Kb[b][t][d] = K(t, t - d), d = 0..bw (row t of the lower band, d = 0 the diagonal), W = bw + 1 doubles per
// row.  After factorisation: d = 0 holds D_t, d >= 1 holds L(t, t - d).
//
// One WARP per matrix (these batches are small: 128 instances per GPU in cfg4), right-looking: a sliding window of
// bw + 1 rows lives in shared memory; per column: reciprocal, scale the column, rank-1 update of the window triangle.
#include "gf_common.cuh"
#include "../../include/gradflow_b200.h"

namespace {

constexpr int BAND_MAX = 64;  // bw + 1 <= 64

// K(i, j) of the full regularised KKT matrix, full indices (variables 0..n-1, constraints n..n+m-1), i >= j in
// band order is NOT assumed here: symmetric lookup.
__device__ __forceinline__ double kkt_entry(int i, int j, int n, const double* __restrict__ Hb,
                                            const double* __restrict__ Jb, const uint8_t* __restrict__ act, double lamb,
                                            double delta) {
    const bool vi = i < n, vj = j < n;
    if (vi && vj) {
        if (act[i] || act[j]) return (i == j) ? 1.0 : 0.0;
        const double h = Hb[(size_t)i * n + j];
        return (i == j) ? h + lamb : h;
    }
    if (!vi && !vj) return (i == j) ? -delta : 0.0;
    const int c = vi ? j - n : i - n, v = vi ? i : j;
    return act[v] ? 0.0 : Jb[(size_t)c * n + v];
}

__global__ void band_assemble_kernel(int n, int m, int bw, const double* __restrict__ H, const double* __restrict__ J,
                                     const uint8_t* __restrict__ active, const int32_t* __restrict__ order,
                                     const double* __restrict__ dt, const double* __restrict__ rho,
                                     double* __restrict__ Kb, GfWork work) {
    const int b = gf_instance(work, blockIdx.y);
    if (b < 0) return;
    const int N = n + m, W = bw + 1;
    const double lamb = 1.0 / dt[b];
    const double delta = lamb / (1.0 + lamb * rho[b]);  // symmetric_step_solver.py:62-66: lamb * fact
    const double* Hb = H + (size_t)b * n * n;
    const double* Jb = J != nullptr ? J + (size_t)b * m * n : nullptr;
    const uint8_t* act = active + (size_t)b * n;
    double* out = Kb + (size_t)b * N * W;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < N * W; e += gridDim.x * blockDim.x) {
        const int t = e / W, d = e - t * W;
        double v = 0.0;
        if (t - d >= 0) v = kkt_entry(order[t], order[t - d], n, Hb, Jb, act, lamb, delta);
        out[e] = v;
    }
}

// In-place banded L D L'.  info[b]: 0 ok, t + 1 zero / non-finite pivot at band position t; nneg[b] = negative pivots.
// W = bw + 1 must be even (rows are moved with 16-byte cp.async); the window has PF extra slots so that the row
// entering at step j is not needed before step j + PF + 1.
constexpr int BAND_PF = 8;

__global__ void __launch_bounds__(32) band_factor_kernel(int N, int bw, double* __restrict__ Kb,
                                                         int32_t* __restrict__ info, int32_t* __restrict__ nneg,
                                                         GfWork work) {
    const int b = gf_instance(work, blockIdx.x);
    if (b < 0) return;
    extern __shared__ double bsm[];
    const int W = bw + 1, S = W + BAND_PF, lane = threadIdx.x;
    double* win = bsm;           // S x W circular window: row r at slot r % S
    double* lv = win + S * W;    // L(r, j) of the current column, r = j + 1 + index
    double* wv = lv + BAND_MAX;  // the same entries unscaled (= L D)
    unsigned short* pairs = reinterpret_cast<unsigned short*>(wv + BAND_MAX);  // (kr << 8) | kc, kc <= kr < bw
    double* Kmat = Kb + (size_t)b * N * W;
    int bad = 0, neg = 0;
    const int npair = bw * (bw + 1) / 2;
    for (int e = lane; e < npair; e += 32) {  // row-major enumeration of the lower triangle
        int kr = (int)((sqrtf(8.0f * e + 1.0f) - 1.0f) * 0.5f);
        if ((kr + 1) * (kr + 2) / 2 <= e) kr++;
        if (kr * (kr + 1) / 2 > e) kr--;
        pairs[e] = (unsigned short)((kr << 8) | (e - kr * (kr + 1) / 2));
    }
    for (int r = 0; r < min(N, S); r++)
        for (int d = lane; d < W; d += 32) win[(r % S) * W + d] = Kmat[(size_t)r * W + d];
    __syncwarp();
    int sj = 0;  // slot of row j
    for (int j = 0; j < N; j++) {
        cp_async_wait<BAND_PF>();  // every row up to j + bw has landed
        __syncwarp();
        const double dj = win[sj * W];
        if (!(isfinite(dj)) || dj == 0.0) { if (bad == 0) bad = j + 1; }
        if (dj < 0.0) neg++;
        double rinv;
        {   // 1 / dj: MUFU seed + cubic step + Newton step (a full division is ~3x longer on this critical path)
            double r0;
            asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(dj));
            const double e = fma(-dj, r0, 1.0);
            r0 = fma(r0, fma(e, e, e), r0);
            rinv = fma(r0, fma(-dj, r0, 1.0), r0);
            if (dj == 0.0) rinv = 0.0;
        }
        const int nb = min(bw, N - 1 - j);  // rows below inside the band
        const int s1 = sj + 1;              // slot of row j + 1 (before wrap-around)
        // column j: row r = j + 1 + k has its entry at d = k + 1
        for (int k = lane; k < nb; k += 32) {
            int sl = s1 + k;
            if (sl >= S) sl -= S;
            const int slot = sl * W + (k + 1);
            const double w = win[slot];
            const double l = w * rinv;
            wv[k] = w;
            lv[k] = l;
            win[slot] = l;
        }
        __syncwarp();
        // window update: K(r, c) -= L(r, j) D_j L(c, j) for j < c <= r <= j + nb; entry (r, c) sits at d = r - c.
        // The pairs are independent, so they are spread over the lanes.
        const int np = nb * (nb + 1) / 2;
        for (int e = lane; e < np; e += 32) {
            const int pr = pairs[e], kr = pr >> 8, kc = pr & 255;
            int sl = s1 + kr;
            if (sl >= S) sl -= S;
            double* ep = win + sl * W + (kr - kc);
            *ep = fma(-lv[kr], wv[kc], *ep);
        }
        __syncwarp();
        // row j is final: write it back; row j + S is fetched into its slot (asynchronously)
        {
            double* src = win + sj * W;
            for (int d = lane; d < W; d += 32) Kmat[(size_t)j * W + d] = src[d];
            __syncwarp();
            const int rn = j + S;
            if (rn < N)
                for (int d2 = lane; d2 < W / 2; d2 += 32) cp_async16(src + 2 * d2, Kmat + (size_t)rn * W + 2 * d2);
            cp_async_commit();
        }
        sj = (sj + 1 == S) ? 0 : sj + 1;
    }
    cp_async_wait<0>();
    if (lane == 0) {
        info[b] = bad;
        nneg[b] = neg;
    }
}

// x = K^{-1} r in band order: L z = r, z /= D, L' x = z.  v[b] (length N) is overwritten.  Both sweeps are written in
// "axpy" form (as soon as z_t is final it is subtracted from the bw entries it touches), so a step has no warp
// reduction on its critical path; the band entries are fetched eight steps ahead into registers.
__global__ void __launch_bounds__(32) band_solve_kernel(int N, int bw, const double* __restrict__ Kb,
                                                        double* __restrict__ v, GfWork work) {
    const int b = gf_instance(work, blockIdx.x);
    if (b < 0) return;
    extern __shared__ double z[];  // N doubles
    const int W = bw + 1, lane = threadIdx.x;
    const double* Kmat = Kb + (size_t)b * N * W;
    double* vb = v + (size_t)b * N;
    for (int t = lane; t < N; t += 32) z[t] = vb[t];
    __syncwarp();
    const int d0 = 1 + lane, d1 = 33 + lane;  // this lane's band offsets (bw <= 63)
    // forward, unit lower: z_{t+d} -= L(t+d, t) z_t  (column t of L: Kmat[t+d][d])
    auto fetch_fwd = [&](int t0, double (&a0)[8], double (&a1)[8]) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const int t = t0 + u;
            a0[u] = (d0 <= bw && t + d0 < N) ? Kmat[(size_t)(t + d0) * W + d0] : 0.0;
            a1[u] = (d1 <= bw && t + d1 < N) ? Kmat[(size_t)(t + d1) * W + d1] : 0.0;
        }
    };
    auto fetch_bwd = [&](int t0, double (&a0)[8], double (&a1)[8]) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const int t = t0 - u;
            a0[u] = (t >= 0 && d0 <= bw && t - d0 >= 0) ? Kmat[(size_t)t * W + d0] : 0.0;
            a1[u] = (t >= 0 && d1 <= bw && t - d1 >= 0) ? Kmat[(size_t)t * W + d1] : 0.0;
        }
    };
    double a0[8], a1[8], n0[8], n1[8];
    fetch_fwd(0, a0, a1);
    for (int t0 = 0; t0 < N; t0 += 8) {
        fetch_fwd(t0 + 8, n0, n1);  // the next eight steps' entries are in flight while these eight are applied
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const int t = t0 + u;
            if (t < N) {
                const double zt = z[t];
                if (d0 <= bw && t + d0 < N) z[t + d0] = fma(-a0[u], zt, z[t + d0]);
                if (d1 <= bw && t + d1 < N) z[t + d1] = fma(-a1[u], zt, z[t + d1]);
                __syncwarp();
            }
        }
#pragma unroll
        for (int u = 0; u < 8; u++) { a0[u] = n0[u]; a1[u] = n1[u]; }
    }
    for (int t = lane; t < N; t += 32) z[t] /= Kmat[(size_t)t * W];
    __syncwarp();
    // backward: x_{t-d} -= L(t, t-d) x_t  (row t of L: Kmat[t][d], contiguous)
    fetch_bwd(N - 1, a0, a1);
    for (int t0 = N - 1; t0 >= 0; t0 -= 8) {
        fetch_bwd(t0 - 8, n0, n1);
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const int t = t0 - u;
            if (t >= 0) {
                const double xt = z[t];
                if (d0 <= bw && t - d0 >= 0) z[t - d0] = fma(-a0[u], xt, z[t - d0]);
                if (d1 <= bw && t - d1 >= 0) z[t - d1] = fma(-a1[u], xt, z[t - d1]);
                __syncwarp();
            }
        }
#pragma unroll
        for (int u = 0; u < 8; u++) { a0[u] = n0[u]; a1[u] = n1[u]; }
    }
    for (int t = lane; t < N; t += 32) vb[t] = z[t];
}

// rhs in the reduced standard order (inactive variables in perm order, then constraints) -> band order of the
// full system (active positions get 0), and back.
__global__ void band_permute_kernel(int n, int m, int ld, const int32_t* __restrict__ perm,
                                    const int32_t* __restrict__ nIv, const int32_t* __restrict__ pos,
                                    double* __restrict__ stdv, double* __restrict__ bandv, int to_band, GfWork work) {
    const int b = gf_instance(work, blockIdx.x);
    if (b < 0) return;
    const int N = n + m, nI = nIv[b];
    double* sv = stdv + (size_t)b * ld;
    double* bv = bandv + (size_t)b * N;
    if (to_band) {
        for (int t = threadIdx.x; t < N; t += blockDim.x) bv[t] = 0.0;
        __syncthreads();
    }
    for (int r = threadIdx.x; r < nI + m; r += blockDim.x) {
        const int full = r < nI ? perm[(size_t)b * n + r] : n + (r - nI);
        const int t = pos[full];
        if (to_band) bv[t] = sv[r];
        else sv[r] = bv[t];
    }
}

}  // namespace

extern "C" int gf_band_assemble(int B, int n, int m, int bw, const double* H, const double* J, const uint8_t* active,
                                const int32_t* order, const double* dt, const double* rho, double* Kband,
                                const int32_t* work, const int32_t* nwork_dev, int nwork, void* stream) {
    if (B <= 0 || n <= 0 || m < 0 || bw < 0 || bw + 1 > BAND_MAX || !H || !active || !order || !dt || !rho || !Kband)
        return GF_ERR_ARG;
    if (m > 0 && !J) return GF_ERR_ARG;
    if (nwork <= 0) return GF_OK;
    if (nwork > 65535) return GF_ERR_UNSUPPORTED;  // the instance index rides in grid.y
    const int total = (n + m) * (bw + 1);
    int gx = (total + 255) / 256;
    if (gx > 64) gx = 64;
    band_assemble_kernel<<<dim3(gx, nwork), 256, 0, (cudaStream_t)stream>>>(n, m, bw, H, J, active, order, dt, rho, Kband,
                                                                            GfWork{work, nwork_dev});
    return gf_launch_status();
}

extern "C" int gf_band_factor(int B, int N, int bw, double* Kband, int32_t* info, int32_t* nneg, const int32_t* work,
                              const int32_t* nwork_dev, int nwork, void* stream) {
    if (B <= 0 || N <= 0 || bw < 0 || bw + 1 > BAND_MAX || ((bw + 1) & 1) || !Kband || !info || !nneg) return GF_ERR_ARG;
    if (nwork <= 0) return GF_OK;
    const int W = bw + 1;
    const size_t smem = ((size_t)(W + BAND_PF) * W + 2 * BAND_MAX) * sizeof(double) +
                        (size_t)(bw * (bw + 1) / 2 + 4) * sizeof(unsigned short);
    if (smem > 48 * 1024) cudaFuncSetAttribute(band_factor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    band_factor_kernel<<<nwork, 32, smem, (cudaStream_t)stream>>>(N, bw, Kband, info, nneg, GfWork{work, nwork_dev});
    return gf_launch_status();
}

extern "C" int gf_band_solve(int B, int N, int bw, const double* Kband, double* v, const int32_t* work,
                             const int32_t* nwork_dev, int nwork, void* stream) {
    if (B <= 0 || N <= 0 || bw < 0 || bw + 1 > BAND_MAX || !Kband || !v) return GF_ERR_ARG;
    if (nwork <= 0) return GF_OK;
    const size_t smem = (size_t)N * sizeof(double);
    if (smem > 200 * 1024) return GF_ERR_UNSUPPORTED;
    if (smem > 48 * 1024) cudaFuncSetAttribute(band_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    band_solve_kernel<<<nwork, 32, smem, (cudaStream_t)stream>>>(N, bw, Kband, v, GfWork{work, nwork_dev});
    return gf_launch_status();
}

extern "C" int gf_band_permute(int B, int n, int m, int ld, const int32_t* perm, const int32_t* nI, const int32_t* pos,
                               double* stdv, double* bandv, int to_band, const int32_t* work, const int32_t* nwork_dev,
                               int nwork, void* stream) {
    if (B <= 0 || n <= 0 || m < 0 || ld < n + m || !perm || !nI || !pos || !stdv || !bandv) return GF_ERR_ARG;
    if (nwork <= 0) return GF_OK;
    band_permute_kernel<<<nwork, 256, 0, (cudaStream_t)stream>>>(n, m, ld, perm, nI, pos, stdv, bandv, to_band,
                                                                 GfWork{work, nwork_dev});
    return gf_launch_status();
}
