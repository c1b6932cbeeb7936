// H_rho = H + rho J'J for the Standard step formulation (iterate.py:103-110 aug_lag_deriv_xx(rho) = lag_hess + rho J'J,
// used by standard_step_solver.py:50-53): batched rank-m update on the FP64 tensor pipe (mma.sync m8n8k4 -> DMMA).
//
// One CTA (4 warps, 2 x 2) per 64 x 64 tile of the n x n result and instance; J[b] (m x n, row-major) is both operands:
// A = J' (A[i][k] = J[k][i]) and B = J, so a k-chunk of 32 rows of J is staged once per operand column block in shared
// memory (pitch 68 doubles: the 8-byte fragment loads of a half warp hit 16 distinct bank pairs) and every warp forms a
// 32 x 32 sub-tile from 4 x 4 DMMA tiles.  The next chunk is fetched into registers while the current one is multiplied.
// Epilogue: out = rho * C + H, un-fused like the expression it replaces (rho * (J'J) + H).
// Bound: FP64 tensor pipe, 2 n^2 m flop per instance (the result is symmetric; both triangles are computed so that every
// store is a coalesced row segment).
#include "gf_common.cuh"
#include "../../include/gradflow_b200.h"

namespace {

constexpr int ST = 64;        // tile
constexpr int SK = 32;        // k-chunk
constexpr int SP = ST + 4;    // shared-memory pitch

__global__ void __launch_bounds__(128, 3) hess_rho_kernel(int n, int m, const double* __restrict__ H,
                                                       const double* __restrict__ J, const double* __restrict__ rho,
                                                       double* __restrict__ out, GfWork work) {
    const int b = gf_instance(work, blockIdx.z);
    if (b < 0) return;
    __shared__ double As[SK][SP];
    __shared__ double Bs[SK][SP];
    const int i0 = blockIdx.y * ST, j0 = blockIdx.x * ST;
    const double* Jb = J + (size_t)b * m * n;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int wi = (wid >> 1) * 32, wj = (wid & 1) * 32;
    double c[4][4][2];
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int q = 0; q < 4; ++q) c[p][q][0] = c[p][q][1] = 0.0;
    // each thread stages 16 elements per operand and chunk: row kr = e / 64 ... (32 x 64 = 2048 = 128 x 16)
    double ra[16], rb[16];
    auto fetch = [&](int k0) {
#pragma unroll
        for (int e = 0; e < 16; ++e) {
            const int idx = tid + 128 * e;
            const int kr = idx >> 6, cc = idx & 63;
            const int k = k0 + kr;
            ra[e] = (k < m && i0 + cc < n) ? __ldg(Jb + (size_t)k * n + i0 + cc) : 0.0;
            rb[e] = (k < m && j0 + cc < n) ? __ldg(Jb + (size_t)k * n + j0 + cc) : 0.0;
        }
    };
    fetch(0);
    for (int k0 = 0; k0 < m; k0 += SK) {
        __syncthreads();
#pragma unroll
        for (int e = 0; e < 16; ++e) {
            const int idx = tid + 128 * e;
            As[idx >> 6][idx & 63] = ra[e];
            Bs[idx >> 6][idx & 63] = rb[e];
        }
        __syncthreads();
        if (k0 + SK < m) fetch(k0 + SK);
#pragma unroll
        for (int kk = 0; kk < SK; kk += 4) {
            double af[4], bf[4];
#pragma unroll
            for (int p = 0; p < 4; ++p) af[p] = As[kk + (lane & 3)][wi + 8 * p + (lane >> 2)];
#pragma unroll
            for (int q = 0; q < 4; ++q) bf[q] = Bs[kk + (lane & 3)][wj + 8 * q + (lane >> 2)];
#pragma unroll
            for (int p = 0; p < 4; ++p)
#pragma unroll
                for (int q = 0; q < 4; ++q) dmma884(c[p][q][0], c[p][q][1], af[p], bf[q]);
        }
    }
    const double r = rho[b];
    const double* Hb = H + (size_t)b * n * n;
    double* ob = out + (size_t)b * n * n;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int i = i0 + wi + 8 * p + (lane >> 2);
        if (i >= n) continue;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int j = j0 + wj + 8 * q + 2 * (lane & 3);
            if (j < n) ob[(size_t)i * n + j] = __dadd_rn(__dmul_rn(r, c[p][q][0]), Hb[(size_t)i * n + j]);
            if (j + 1 < n) ob[(size_t)i * n + j + 1] = __dadd_rn(__dmul_rn(r, c[p][q][1]), Hb[(size_t)i * n + j + 1]);
        }
    }
}

}  // namespace

extern "C" int gf_hess_rho(int B, int n, int m, const double* H, const double* J, const double* rho, double* out,
                           const int32_t* work, const int32_t* nwork_dev, int nwork, void* stream) {
    if (B <= 0 || n <= 0 || m <= 0 || !H || !J || !rho || !out || nwork < 0) return GF_ERR_ARG;
    if (nwork == 0) return GF_OK;
    if (nwork > 65535) return GF_ERR_UNSUPPORTED;  // the instance index rides in grid.z
    const int tiles = (n + ST - 1) / ST;
    dim3 grid(tiles, tiles, nwork);
    hess_rho_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(n, m, H, J, rho, out, GfWork{work, nwork_dev});
    return gf_launch_status();
}
