// Batched dense FP64 LDL' (no pivoting) for symmetric quasi-definite KKT matrices, trailing update on the
// FP64 tensor pipe (mma.sync m8n8k4 f64 -> SASS DMMA.8x8x4), + substitution (a13, a14, inertia contract).
//
// Replaces pygradflow/linear_solver/lu_solver.py:9-21 on the Symmetric step solver's matrix
// K = [[H_II + lamb I, J_I'], [J_I, -delta I]] (symmetric_step_solver.py:49-77), which is quasi-definite
// whenever H_II + lamb I > 0, so L D L' exists for every symmetric permutation and is stable without
// pivoting; the signs of D give the inertia that ma57_solver.py:76-79 / mumps_solver.py:81-82 /
// ssids_solver.py:22-23 report (num_neg_eigvals).
//
// Layout: K[b] ld x ld row-major, only the lower triangle is read/written; the matrix is padded with an
// identity block up to Np = roundup(N_b, 64) (gf_kkt_assemble does this) so no kernel has edge tiles
// inside a 64-block.  d is kept on the diagonal of K and, contiguously, in dvec[b].
//
// Left-looking by 64-wide block columns, three batched launches per block column k:
//   update(k)  A[i, k] -= sum_{p<k} (L[i,p] D_p) L[k,p]'   128x64 tiles, DMMA, cp.async 3-stage pipeline
//   diag(k)    A[k, k] = L_kk D_k L_kk'                    64x64 in shared memory
//   trsm(k)    L[i, k] = A[i, k] L_kk^{-T} D_k^{-1}        thread-per-row substitution (full-rate DFMA)
// Each matrix entry is read/written O(1) times (left-looking), operands stream once per block column:
// ~ 8 N^3 / (6*64) bytes per matrix -> the update is bound by the FP64 pipe, not HBM.
#include "gf_common.cuh"
#include "../../include/gradflow_b200.h"

namespace {

constexpr int NB = 64;

__device__ __forceinline__ int padded_order(const int32_t* Nvec, int Nfixed, int b, int ld) {
    const int N = Nvec != nullptr ? Nvec[b] : Nfixed;
    const int Np = ((N + NB - 1) / NB) * NB;
    return Np < ld ? Np : ld;
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ldlt_diag_kernel(int ld, const int32_t* __restrict__ Nvec, int Nfixed, int k,
                                                        double* __restrict__ K, double* __restrict__ dvec,
                                                        int32_t* __restrict__ info, int32_t* __restrict__ nneg,
                                                        const int32_t* __restrict__ npos_expected, GfWork work) {
    const int b = gf_instance(work, blockIdx.x);
    if (b < 0) return;
    const int Np = padded_order(Nvec, Nfixed, b, ld);
    const int j0 = k * NB;
    if (j0 >= Np) return;
    __shared__ double S[NB][NB + 1];
    __shared__ int s_bad, s_neg, s_sign;
    double* Kb = K + (size_t)b * ld * ld;
    if (threadIdx.x == 0) { s_bad = 0x7fffffff; s_neg = 0; s_sign = 0; }
    for (int e = threadIdx.x; e < NB * NB; e += blockDim.x) {
        const int i = e >> 6, c = e & 63;
        S[i][c] = (c <= i) ? Kb[(size_t)(j0 + i) * ld + j0 + c] : 0.0;
    }
    __syncthreads();
    const int i = threadIdx.x & 63, cg = threadIdx.x >> 6;
    // right-looking on W = L D kept in place; scaled to L at the end
    for (int j = 0; j < NB - 1; j++) {
        const double d = S[j][j];
        if (i > j && d != 0.0) {
            const double li = S[i][j] / d;
            for (int c = j + 1 + ((cg - (j + 1)) & 3); c <= i; c += 4) S[i][c] -= li * S[c][j];
        }
        __syncthreads();
    }
    if (threadIdx.x < NB) {
        const double d = S[threadIdx.x][threadIdx.x];
        if (!(isfinite(d)) || d == 0.0) atomicMin(&s_bad, (int)threadIdx.x + 1);
        const int N = Nvec != nullptr ? Nvec[b] : Nfixed;
        const int j = j0 + (int)threadIdx.x;
        if (d < 0.0 && j < N) atomicAdd(&s_neg, 1);
        // quasi-definite sign pattern: the first npos pivots positive, the remaining (up to N) negative
        if (npos_expected != nullptr && j < N && ((j < npos_expected[b]) != (d > 0.0))) s_sign = 1;
        dvec[(size_t)b * ld + j0 + threadIdx.x] = d;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < NB * NB; e += blockDim.x) {
        const int r = e >> 6, c = e & 63;
        if (c < r) Kb[(size_t)(j0 + r) * ld + j0 + c] = S[r][c] / S[c][c];
        else if (c == r) Kb[(size_t)(j0 + r) * ld + j0 + c] = S[r][r];
    }
    if (threadIdx.x == 0) {
        int bad = 0;
        if (s_bad != 0x7fffffff) bad = j0 + s_bad;        // zero / non-finite pivot at column `bad`
        else if (s_sign) bad = GF_INFO_NOT_QUASIDEFINITE;  // wrong pivot sign: unpivoted LDL' not trusted
        if (k == 0) { nneg[b] = s_neg; info[b] = bad; }
        else {
            nneg[b] += s_neg;
            if (bad != 0 && info[b] == 0) info[b] = bad;
        }
    }
}

// ------------------------------------------------------------------------------------------------
constexpr int TR_ROWS = 128;

__global__ void __launch_bounds__(TR_ROWS) ldlt_trsm_kernel(int ld, const int32_t* __restrict__ Nvec, int Nfixed,
                                                            int k, double* __restrict__ K,
                                                            const double* __restrict__ dvec, GfWork work) {
    const int b = gf_instance(work, blockIdx.y);
    if (b < 0) return;
    const int Np = padded_order(Nvec, Nfixed, b, ld);
    const int j0 = k * NB;
    const int i0 = j0 + NB + blockIdx.x * TR_ROWS;
    if (i0 >= Np) return;
    const int nrows = min(TR_ROWS, Np - i0);
    extern __shared__ double sm[];
    double* Lt = sm;                     // Lt[p*64 + c] = L_kk[c][p] (c > p)
    double* dk = Lt + NB * NB;           // 64
    double* Ws = dk + NB;                // TR_ROWS x 65
    double* Kb = K + (size_t)b * ld * ld;
    for (int e = threadIdx.x; e < NB * NB; e += blockDim.x) {
        const int c = e >> 6, p = e & 63;   // coalesced read of storage row j0 + c
        if (p < c) Lt[p * NB + c] = Kb[(size_t)(j0 + c) * ld + j0 + p];
    }
    if (threadIdx.x < NB) dk[threadIdx.x] = dvec[(size_t)b * ld + j0 + threadIdx.x];
    for (int e = threadIdx.x; e < nrows * NB; e += blockDim.x) {
        const int r = e >> 6, c = e & 63;
        Ws[r * (NB + 1) + c] = Kb[(size_t)(i0 + r) * ld + j0 + c];
    }
    __syncthreads();
    const int r = threadIdx.x;
    if (r < nrows) {
        double* w = Ws + r * (NB + 1);
#pragma unroll 1
        for (int cb = 0; cb < NB; cb += 8) {
            double acc[8];
#pragma unroll
            for (int u = 0; u < 8; u++) acc[u] = w[cb + u];
            for (int p = 0; p < cb; p++) {
                const double wp = w[p];
                const double2* l2 = reinterpret_cast<const double2*>(Lt + p * NB + cb);
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const double2 l = l2[u];
                    acc[2 * u] = fma(-wp, l.x, acc[2 * u]);
                    acc[2 * u + 1] = fma(-wp, l.y, acc[2 * u + 1]);
                }
            }
#pragma unroll
            for (int u = 1; u < 8; u++)
#pragma unroll
                for (int q = 0; q < u; q++) acc[u] = fma(-acc[q], Lt[(cb + q) * NB + cb + u], acc[u]);
#pragma unroll
            for (int u = 0; u < 8; u++) w[cb + u] = acc[u];
        }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < nrows * NB; e += blockDim.x) {
        const int rr = e >> 6, c = e & 63;
        Kb[(size_t)(i0 + rr) * ld + j0 + c] = Ws[rr * (NB + 1) + c] / dk[c];
    }
}

// ------------------------------------------------------------------------------------------------
constexpr int TM = 128, TN = 64, KC = 16, STAGES = 3, SP = KC + 4;  // SP: smem row pitch (bank-conflict free)

__global__ void __launch_bounds__(256, 2) ldlt_update_kernel(int ld, const int32_t* __restrict__ Nvec, int Nfixed,
                                                             int k, double* __restrict__ K,
                                                             const double* __restrict__ dvec, GfWork work) {
    const int b = gf_instance(work, blockIdx.y);
    if (b < 0) return;
    const int Np = padded_order(Nvec, Nfixed, b, ld);
    const int j0 = k * NB;
    const int i0 = j0 + blockIdx.x * TM;
    if (i0 >= Np) return;
    extern __shared__ double sm[];
    double* As = sm;                          // STAGES x TM x SP
    double* Bs = sm + STAGES * TM * SP;       // STAGES x TN x SP
    double* Kb = K + (size_t)b * ld * ld;
    const double* db = dvec + (size_t)b * ld;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int wm = wid >> 1, wn = wid & 1;    // 4 x 2 warps, 32 x 32 each
    const int g = lane >> 2, q = lane & 3;
    const int nchunks = j0 / KC;

    auto load_stage = [&](int chunk, int stage) {
        const int p0 = chunk * KC;
        double* as = As + stage * TM * SP;
        double* bs = Bs + stage * TN * SP;
#pragma unroll
        for (int t = 0; t < 4; t++) {
            const int piece = tid + t * 256;          // 1024 pieces of 16 B: row = piece / 8, part = piece % 8
            const int r = piece >> 3, part = piece & 7;
            double* dst = as + r * SP + part * 2;
            if (i0 + r < Np) cp_async16(dst, Kb + (size_t)(i0 + r) * ld + p0 + part * 2);
            else { dst[0] = 0.0; dst[1] = 0.0; }
        }
#pragma unroll
        for (int t = 0; t < 2; t++) {
            const int piece = tid + t * 256;          // 512 pieces
            const int r = piece >> 3, part = piece & 7;
            cp_async16(bs + r * SP + part * 2, Kb + (size_t)(j0 + r) * ld + p0 + part * 2);
        }
    };

    // accumulators start from the current A tile
    double acc[4][4][2];
#pragma unroll
    for (int mi = 0; mi < 4; mi++) {
        const int row = i0 + wm * 32 + mi * 8 + g;
#pragma unroll
        for (int ni = 0; ni < 4; ni++) {
            const int col = j0 + wn * 32 + ni * 8 + 2 * q;
            if (row < Np) {
                const double2 v = *reinterpret_cast<const double2*>(Kb + (size_t)row * ld + col);
                acc[mi][ni][0] = v.x;
                acc[mi][ni][1] = v.y;
            } else {
                acc[mi][ni][0] = 0.0;
                acc[mi][ni][1] = 0.0;
            }
        }
    }

#pragma unroll
    for (int s = 0; s < STAGES - 1; s++) {
        if (s < nchunks) load_stage(s, s);
        cp_async_commit();
    }
    for (int c = 0; c < nchunks; c++) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        const int nxt = c + STAGES - 1;
        if (nxt < nchunks) load_stage(nxt, nxt % STAGES);
        cp_async_commit();
        const double* as = As + (c % STAGES) * TM * SP + (wm * 32 + g) * SP + q;
        const double* bs = Bs + (c % STAGES) * TN * SP + (wn * 32 + g) * SP + q;
        const double* dp = db + c * KC + q;
#pragma unroll
        for (int kk = 0; kk < KC; kk += 4) {
            const double nd = -__ldg(dp + kk);
            double a[4], bf[4];
#pragma unroll
            for (int mi = 0; mi < 4; mi++) a[mi] = as[mi * 8 * SP + kk] * nd;
#pragma unroll
            for (int ni = 0; ni < 4; ni++) bf[ni] = bs[ni * 8 * SP + kk];
#pragma unroll
            for (int mi = 0; mi < 4; mi++)
#pragma unroll
                for (int ni = 0; ni < 4; ni++) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], bf[ni]);
        }
    }
    cp_async_wait<0>();
#pragma unroll
    for (int mi = 0; mi < 4; mi++) {
        const int row = i0 + wm * 32 + mi * 8 + g;
        if (row < Np) {
#pragma unroll
            for (int ni = 0; ni < 4; ni++) {
                const int col = j0 + wn * 32 + ni * 8 + 2 * q;
                *reinterpret_cast<double2*>(Kb + (size_t)row * ld + col) = make_double2(acc[mi][ni][0], acc[mi][ni][1]);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// x = K^{-1} r via L z = r, z /= d, L' x = z.  rhs[b] (length >= N) is overwritten.
__global__ void __launch_bounds__(256) ldlt_solve_kernel(int ld, const int32_t* __restrict__ Nvec, int Nfixed,
                                                         const double* __restrict__ K, double* __restrict__ rhs,
                                                         int ldr, GfWork work) {
    const int b = gf_instance(work, blockIdx.x);
    if (b < 0) return;
    const int N = Nvec != nullptr ? Nvec[b] : Nfixed;
    if (N <= 0) return;
    extern __shared__ double v[];
    __shared__ double Tb[32][33];
    __shared__ double part[32];
    const double* Kb = K + (size_t)b * ld * ld;
    double* rb = rhs + (size_t)b * ldr;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int i = threadIdx.x; i < N; i += blockDim.x) v[i] = rb[i];
    __syncthreads();
    // forward (unit lower, row-contiguous dots), then divide by d
    for (int j0 = 0; j0 < N; j0 += 32) {
        const int jb = min(32, N - j0);
        for (int e = threadIdx.x; e < 32 * 32; e += blockDim.x) {
            const int jj = e >> 5, ii = e & 31;
            Tb[jj][ii] = (jj < jb && ii <= jj) ? Kb[(size_t)(j0 + jj) * ld + j0 + ii] : 0.0;
        }
        for (int jj = wid; jj < jb; jj += nw) {
            const double* row = Kb + (size_t)(j0 + jj) * ld;
            double acc = 0.0;
            for (int i = lane; i < j0; i += 32) acc += __ldg(row + i) * v[i];
            acc = warp_sum(acc);
            if (lane == 0) part[jj] = acc;
        }
        __syncthreads();
        if (wid == 0) {
            double s = (lane < jb) ? v[j0 + lane] - part[lane] : 0.0;
            for (int ii = 0; ii < jb; ii++) {
                const double w = __shfl_sync(0xffffffffu, s, ii);
                if (lane > ii && lane < jb) s -= Tb[lane][ii] * w;
            }
            if (lane < jb) v[j0 + lane] = s;  // still z (undivided); d applied below
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < N; i += blockDim.x) v[i] /= Kb[(size_t)i * ld + i];
    __syncthreads();
    // backward: L' x = z, axpy form over storage rows
    const int nblk = (N + 31) / 32;
    for (int kb = nblk - 1; kb >= 0; kb--) {
        const int j0 = kb * 32, jb = min(32, N - j0);
        for (int e = threadIdx.x; e < 32 * 32; e += blockDim.x) {
            const int ii = e >> 5, jj = e & 31;
            Tb[ii][jj] = (ii < jb && jj < ii) ? Kb[(size_t)(j0 + ii) * ld + j0 + jj] : 0.0;
        }
        __syncthreads();
        if (wid == 0) {
            double s = (lane < jb) ? v[j0 + lane] : 0.0;
            for (int ii = jb - 1; ii >= 0; ii--) {
                const double xi = __shfl_sync(0xffffffffu, s, ii);
                if (lane < ii) s -= Tb[ii][lane] * xi;
            }
            if (lane < jb) v[j0 + lane] = s;
        }
        __syncthreads();
        for (int j = threadIdx.x; j < j0; j += blockDim.x) {
            double acc = 0.0;
            for (int ii = 0; ii < jb; ii++) acc += __ldg(Kb + (size_t)(j0 + ii) * ld + j) * v[j0 + ii];
            v[j] -= acc;
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < N; i += blockDim.x) rb[i] = v[i];
}

}  // namespace

extern "C" int gf_ldlt_factor(int B, int ld, int Nmax, const int32_t* Nvec, double* K, double* dvec, int32_t* info,
                              int32_t* nneg, const int32_t* npos_expected, const int32_t* work, const int32_t* nwork_dev, int nwork, void* stream) {
    if (B <= 0 || ld <= 0 || (ld % NB) != 0 || Nmax < 0 || Nmax > ld || !K || !dvec || !info || !nneg)
        return GF_ERR_ARG;
    if (nwork <= 0 || Nmax == 0) return GF_OK;
    cudaStream_t s = (cudaStream_t)stream;
    GfWork w{work, nwork_dev};
    const int Np = ((Nmax + NB - 1) / NB) * NB;
    const int nblk = Np / NB;
    const size_t smem_trsm = (size_t)(NB * NB + NB + TR_ROWS * (NB + 1)) * sizeof(double);
    const size_t smem_upd = (size_t)STAGES * (TM + TN) * SP * sizeof(double);
    cudaFuncSetAttribute(ldlt_trsm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_trsm);
    cudaFuncSetAttribute(ldlt_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_upd);
    for (int k = 0; k < nblk; k++) {
        const int j0 = k * NB;
        if (k > 0) {
            dim3 grid((Np - j0 + TM - 1) / TM, nwork);
            ldlt_update_kernel<<<grid, 256, smem_upd, s>>>(ld, Nvec, Nmax, k, K, dvec, w);
        }
        ldlt_diag_kernel<<<nwork, 256, 0, s>>>(ld, Nvec, Nmax, k, K, dvec, info, nneg, npos_expected, w);
        if (j0 + NB < Np) {
            dim3 grid((Np - j0 - NB + TR_ROWS - 1) / TR_ROWS, nwork);
            ldlt_trsm_kernel<<<grid, TR_ROWS, smem_trsm, s>>>(ld, Nvec, Nmax, k, K, dvec, w);
        }
    }
    return gf_launch_status();
}

extern "C" int gf_ldlt_solve(int B, int ld, int Nmax, const int32_t* Nvec, const double* K, double* rhs, int ldr,
                             const int32_t* work, const int32_t* nwork_dev, int nwork, void* stream) {
    if (B <= 0 || ld <= 0 || Nmax < 0 || Nmax > ld || ldr < Nmax || !K || !rhs) return GF_ERR_ARG;
    if (nwork <= 0 || Nmax == 0) return GF_OK;
    const size_t smem = (size_t)(Nmax + 1) * sizeof(double);
    if (smem > 200 * 1024) return GF_ERR_UNSUPPORTED;
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(ldlt_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    ldlt_solve_kernel<<<nwork, 256, smem, (cudaStream_t)stream>>>(ld, Nvec, Nmax, K, rhs, ldr,
                                                                  GfWork{work, nwork_dev});
    return gf_launch_status();
}
