// Batched dense FP64 LU with partial pivoting + forward/back substitution (a13, a14).
//
// Replaces pygradflow/linear_solver/lu_solver.py:9-21 (scipy.sparse.linalg.splu / SuperLU gstrf, gstrs).
//
// Storage: K[b] is ld x ld row-major, the matrix of order N_b sits in its top-left corner.  The kernels
// factor the column-major view M = K' (M(i,j) = K[j*ld + i]) with LAPACK-style row interchanges,
// P M = L U, so the pivot search walks a contiguous storage row (coalesced).  Hence
//   K x = r   <=>  U' w = r, L' v = w, x = P' v        (trans = 0; both sweeps are row-contiguous dots)
//   K' x = r  <=>  L z = P r, U x = z                  (trans = 1; row-contiguous axpys)
// For the symmetric KKT matrix of the Symmetric step solver M = K.
// Factor layout: the pivot sequence `piv` is LAPACK getrf's, but the interchanges of a 32-column block are NOT
// applied to the blocks of L left of it (as LINPACK's dgefa does column by column, here block by block): every
// interchange of a strided storage column costs a 32-byte sector per element, and the left part is half of them.
// gf_lu_solve applies the interchanges block by block accordingly; the factors are only meaningful to it.
//
//   lu_warp_kernel     N <= 32: one warp per matrix, matrix resident in registers, pivoting by warp shuffles.
//   lu_rows_kernel     N <= 64: two warps per matrix, rows resident in registers, pivot row through shared memory
//                      (cfg2, n = 64).
//   lu_smem_kernel     N <= ~110: whole matrix resident in shared memory, one CTA per matrix.
//   lu_regpanel_kernel larger N, per block column: pivoted panel resident in registers (thread per row), the
//                      interchanges of the other storage rows and U12; then lu_update_kernel, the right-looking
//                      trailing update on the FP64 tensor pipe.  (lu_panel_kernel: the shared-memory panel, kept for
//                      more than 2048 rows, 1024 < rows <= 1700, and as the single-launch fallback-list path.)
//                      Robust general path; the throughput path for quasi-definite K is the LDL' in gf_ldlt.cu.
//   lu_solve_kernel  blocked substitution, factors streamed once (HBM-bound).
#include <cstdlib>
#include "gf_common.cuh"
#include <cstdlib>
#include <mutex>
#include <vector>
#include <cstdio>
#include "../../include/gradflow_b200.h"

namespace {

#define GF_LU_TINY 1e-290

struct MaxLoc {
    double v;
    int i;
};

__device__ __forceinline__ MaxLoc maxloc_combine(MaxLoc a, MaxLoc b) {
    // larger |value| wins, ties go to the smaller index (LAPACK idamax); NaN never wins over a number
    if (b.v > a.v || (b.v == a.v && b.i < a.i)) return b;
    return a;
}

__device__ __forceinline__ MaxLoc block_maxloc(MaxLoc m, MaxLoc* scratch) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        MaxLoc other;
        other.v = __shfl_xor_sync(0xffffffffu, m.v, o);
        other.i = __shfl_xor_sync(0xffffffffu, m.i, o);
        m = maxloc_combine(m, other);
    }
    __syncthreads();
    if (lane == 0) scratch[wid] = m;
    __syncthreads();
    MaxLoc r = scratch[0];
    for (int w = 1; w < nw; w++) r = maxloc_combine(r, scratch[w]);
    return r;
}

// absmax = |pivot| as found by the search (-1 when the whole column was NaN)
__device__ __forceinline__ void record_pivot(double absmax, int col, int32_t* info) {
    if (*info == 0) {
        if (absmax < 0.0 || isinf(absmax)) *info = -1;
        else if (absmax < GF_LU_TINY) *info = col + 1;  // exactly singular (lu_solver.py:15-17), or within 1e-290 of it
    }
}

// Warp arg-max of |value| for the pivot search, on the integer pipe: a non-negative double orders like its bit
// pattern, so the maximum is found with three redux.sync (high word, low word among the high-word winners, smallest
// position among the value winners) instead of five dependent rounds of 64-bit shuffles and FP64 compares -- the
// search sits on the critical path of every pivot column.  v < 0 (with idx == none) marks an excluded lane; ties go
// to the smaller position; all lanes excluded -> bv = -1, bi = none.
__device__ __forceinline__ void warp_pivot_search(double v, int idx, int none, double& bv, int& bi) {
    const unsigned long long key = idx == none ? 0ull : (unsigned long long)__double_as_longlong(v) + 1ull;
    const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
    const unsigned mhi = __reduce_max_sync(0xffffffffu, hi);
    const unsigned mlo = __reduce_max_sync(0xffffffffu, hi == mhi ? lo : 0u);
    const bool ismax = hi == mhi && lo == mlo;
    bi = (int)__reduce_min_sync(0xffffffffu, ismax ? (unsigned)idx : (unsigned)none);
    const unsigned long long mk = ((unsigned long long)mhi << 32) | mlo;
    bv = mk == 0ull ? -1.0 : __longlong_as_double((long long)(mk - 1ull));
}

// Quotient by the pivot in the register-resident factorisation kernels: reciprocal (MUFU seed + two Newton steps) times
// the entry, <= 2 ulp, as LAPACK's getf2 scales by the rounded reciprocal -- the IEEE division subroutine costs ~370
// cycles on the column-to-column critical path (one system of order 64: 65 -> 53 us).  The seed flushes denormals, so
// a pivot below GF_LU_TINY in magnitude is treated like an exact zero (record_pivot reports the column, no elimination
// with it); a range check that falls back to the division was measured to cost the whole gain.
__device__ __forceinline__ double lu_pivot_div(double a, double d) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    double e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    return a * r;
}

// ------------------------------------------------------------------------------------------------
__global__ void lu_smem_kernel(int ld, const int32_t* __restrict__ Nvec, int Nfixed, double* __restrict__ K,
                               int32_t* __restrict__ piv, int32_t* __restrict__ info, GfWork work) {
    const int b = gf_instance(work, blockIdx.x);
    if (b < 0) return;
    const int N = Nvec != nullptr ? Nvec[b] : Nfixed;
    extern __shared__ double S[];
    __shared__ MaxLoc scratch[32];
    __shared__ int32_t sinfo;
    const int pitch = N | 1;
    double* Kb = K + (size_t)b * ld * ld;
    int32_t* pb = piv + (size_t)b * ld;
    if (threadIdx.x == 0) sinfo = 0;
    for (int e = threadIdx.x; e < N * N; e += blockDim.x) {
        const int c = e / N, i = e - c * N;
        S[c * pitch + i] = Kb[(size_t)c * ld + i];
    }
    __syncthreads();
    for (int j = 0; j < N; j++) {
        MaxLoc best{-1.0, N};
        for (int i = j + threadIdx.x; i < N; i += blockDim.x) {
            const double a = fabs(S[j * pitch + i]);
            if (a > best.v) { best.v = a; best.i = i; }
        }
        best = block_maxloc(best, scratch);
        const int p = best.i < N ? best.i : j;
        if (threadIdx.x == 0) {
            pb[j] = p;
            record_pivot(best.v, j, &sinfo);
        }
        if (p != j) {
            for (int c = (j & ~63) + threadIdx.x; c < N; c += blockDim.x) {  // not the earlier 64-column super-blocks
                const double t = S[c * pitch + j];
                S[c * pitch + j] = S[c * pitch + p];
                S[c * pitch + p] = t;
            }
        }
        __syncthreads();
        const double pv = S[j * pitch + j];
        if (pv != 0.0) {
            for (int i = j + 1 + threadIdx.x; i < N; i += blockDim.x) {
                const double l = S[j * pitch + i] / pv;
                S[j * pitch + i] = l;
                for (int c = j + 1; c < N; c++) S[c * pitch + i] -= l * S[c * pitch + j];
            }
        }
        __syncthreads();
    }
    for (int e = threadIdx.x; e < N * N; e += blockDim.x) {
        const int c = e / N, i = e - c * N;
        Kb[(size_t)c * ld + i] = S[c * pitch + i];
    }
    if (threadIdx.x == 0) info[b] = sinfo;
}

// ------------------------------------------------------------------------------------------------
// N <= 32: one WARP per matrix, the matrix resident in registers (lane i holds row i of M, 32 doubles), eight
// matrices per CTA, no shared memory and no block barriers.  Pivot search = warp arg-max by shuffles (ties to the
// smaller row, NaN never wins: same rule as block_maxloc); rows are not moved between lanes -- every lane tracks
// the logical position of its row and the factors are written back permuted; the pivot row is broadcast by
// shuffles.  Same arithmetic (division by the pivot, fused multiply-subtract) as lu_smem_kernel.
// The column loop is a real loop over a ROTATING register file: the current column is always a[0]; after the
// elimination step every lane stages its (final) entry of that column in shared memory and shifts its registers
// down by one, so register indices stay static while the code stays a compact loop (a fully unrolled N = 32
// elimination is ~100 KB of straight-line code executed once per warp: instruction-fetch bound).
template <int NMAX>
__global__ void __launch_bounds__(256) lu_warp_kernel(int ld, const int32_t* __restrict__ Nvec, int Nfixed,
                                                      double* __restrict__ K, int32_t* __restrict__ piv,
                                                      int32_t* __restrict__ info, GfWork work, int nwork) {
    extern __shared__ double wstage[];  // 8 warps x NMAX columns x 32 lanes
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slot = blockIdx.x * 8 + wid;
    if (slot >= nwork) return;
    const int b = gf_instance(work, slot);
    if (b < 0) return;
    const int N = Nvec != nullptr ? Nvec[b] : Nfixed;
    double* Kb = K + (size_t)b * ld * ld;
    int32_t* pb = piv + (size_t)b * ld;
    double* stage = wstage + (size_t)wid * NMAX * 32;
    const bool row = lane < N;
    double a[NMAX];
#pragma unroll
    for (int c = 0; c < NMAX; c++) a[c] = (row && c < N) ? Kb[(size_t)c * ld + lane] : 0.0;
    int pos = lane;      // logical row of PM held by this lane (rows never change lanes)
    int32_t sinfo = 0;
    constexpr int NONE = 1 << 20;
    for (int j = 0; j < N; j++) {
        double v = fabs(a[0]);
        int idx = pos;
        if (!(row && pos >= j && v >= 0.0)) { v = -1.0; idx = NONE; }  // excluded rows and NaN entries
        warp_pivot_search(v, idx, NONE, v, idx);
        const int p = idx < NONE ? idx : j;
        record_pivot(v, j, &sinfo);
        const int lp = __ffs(__ballot_sync(0xffffffffu, pos == p)) - 1;  // lane that holds the pivot row
        const int lj = __ffs(__ballot_sync(0xffffffffu, pos == j)) - 1;
        if (lane == lp) pos = j;
        else if (lane == lj) pos = p;
        if (lane == 0) pb[j] = p;
        const double pv = __shfl_sync(0xffffffffu, a[0], lp);
        const bool below = row && pos > j && fabs(pv) >= GF_LU_TINY;
        double l = 0.0;
        if (below) {
            l = a[0] / pv;  // IEEE division here: the smallest systems are the parity tests' reference point
            a[0] = l;
        }
        stage[j * 32 + lane] = a[0];  // final: L entry below the pivot, U entry on and above it
#pragma unroll
        for (int c = 1; c < NMAX; c++) {
            const double pc = __shfl_sync(0xffffffffu, a[c], lp);
            a[c - 1] = below ? fma(-l, pc, a[c]) : a[c];
        }
        a[NMAX - 1] = 0.0;
    }
    __syncwarp();
    if (row)
        for (int c = 0; c < N; c++) Kb[(size_t)c * ld + pos] = stage[c * 32 + lane];
    if (lane == 0) info[b] = sinfo;
}

// 32 < N <= 64: one CTA of two warps per matrix, thread t owns row t in registers, the pivot row travels through
// shared memory (one staged candidate row per warp, one block barrier per column) instead of N shuffles per column.
// Columns are processed in groups of eight with static register indices; after a group the register file is shifted
// down by eight, so the code stays a compact loop (a full unroll is 64 x the column body) at 14 register moves per
// column.  Factor layout as everywhere: the interchanges of the second 32-column block are not applied to the first
// one (pos32).
template <int NMAX, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) lu_rows_kernel(int ld, const int32_t* __restrict__ Nvec, int Nfixed,
                                                             double* __restrict__ K, int32_t* __restrict__ piv,
                                                             int32_t* __restrict__ info, GfWork work, int nwork) {
    constexpr int T = WARPS * 32;
    constexpr int NONE = 1 << 20;
    constexpr int G = 8;
    static_assert(NMAX % G == 0, "group size");
    extern __shared__ __align__(16) double rsm[];
    double* stage = rsm;                       // NMAX x T
    double* wrow = rsm + NMAX * T;             // 2 x WARPS x NMAX
    __shared__ double wval[2][WARPS];
    __shared__ int wpos[2][WARPS];
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
#pragma unroll 1
    for (int wi = blockIdx.x; wi < nwork; wi += gridDim.x) {
        const int b = gf_instance(work, wi);
        if (b < 0) return;
        const int N = Nvec != nullptr ? Nvec[b] : Nfixed;
        double* Kb = K + (size_t)b * ld * ld;
        int32_t* pb = piv + (size_t)b * ld;
        const bool row = t < N;
        double a[NMAX];
#pragma unroll
        for (int c = 0; c < NMAX; c++) a[c] = (row && c < N) ? Kb[(size_t)c * ld + t] : 0.0;
        int pos = t, pos32 = t;
        int32_t sinfo = 0;
#pragma unroll 1
        for (int j0 = 0; j0 < N; j0 += G) {
            if (j0 == 32) pos32 = pos;
#pragma unroll
            for (int jj = 0; jj < G; jj++) {
                const int j = j0 + jj;
                if (j < N) {  // uniform
                    double v = fabs(a[jj]);
                    int idx = pos;
                    if (!(row && pos >= j && v >= 0.0)) { v = -1.0; idx = NONE; }
                    double bv;
                    int bi;
                    warp_pivot_search(v, idx, NONE, bv, bi);
                    const int buf = jj & 1;
                    double2* wr = reinterpret_cast<double2*>(wrow + (size_t)(buf * WARPS + wid) * NMAX);
                    if (bi == NONE) {
                        if (lane == 0) { wval[buf][wid] = -1.0; wpos[buf][wid] = NONE; }
                    } else if (idx == bi) {  // positions are unique: exactly one lane stages its row
                        wval[buf][wid] = bv;
                        wpos[buf][wid] = bi;
#pragma unroll
                        for (int c = 0; c < NMAX; c += 2) wr[c / 2] = make_double2(a[c], a[c + 1]);
                    }
                    __syncthreads();
                    bv = wval[buf][0];
                    bi = wpos[buf][0];
                    int bw = 0;
#pragma unroll
                    for (int w2 = 1; w2 < WARPS; w2++) {
                        const double ov = wval[buf][w2];
                        const int oi = wpos[buf][w2];
                        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; bw = w2; }
                    }
                    const int p = bi < NONE ? bi : j;
                    record_pivot(bv, j, &sinfo);
                    if (t == 0) pb[j] = p;
                    if (pos == p) pos = j;
                    else if (pos == j) pos = p;
                    const double2* pr = reinterpret_cast<const double2*>(wrow + (size_t)(buf * WARPS + bw) * NMAX);
                    const double pv = bi < NONE ? reinterpret_cast<const double*>(pr)[jj] : 0.0;
                    const bool below = row && pos > j && fabs(pv) >= GF_LU_TINY;
                    if (below) {
                        const double l = lu_pivot_div(a[jj], pv);
                        a[jj] = l;
                        // the pivot row in 16-byte loads, eight entries at a time (entries <= jj are dropped)
#pragma unroll
                        for (int c8 = 0; c8 < NMAX; c8 += 8) {
                            double q[8];
#pragma unroll
                            for (int e = 0; e < 8; e += 2) {
                                const double2 v2 = pr[(c8 + e) / 2];
                                q[e] = v2.x;
                                q[e + 1] = v2.y;
                            }
#pragma unroll
                            for (int e = 0; e < 8; e++)
                                if (c8 + e > jj) a[c8 + e] = fma(-l, q[e], a[c8 + e]);
                        }
                    }
                    stage[j * T + t] = a[jj];  // final: L entry below the pivot, U entry on and above it
                }
            }
#pragma unroll
            for (int c = 0; c < NMAX - G; c++) a[c] = a[c + G];
#pragma unroll
            for (int c = NMAX - G; c < NMAX; c++) a[c] = 0.0;
        }
        if (N <= 32) pos32 = pos;
        if (row) {
            for (int c = 0; c < N; c++) Kb[(size_t)c * ld + (c < 32 ? pos32 : pos)] = stage[c * T + t];
        }
        if (t == 0) info[b] = sinfo;
        __syncthreads();
    }
}

template <int NMAX>
int launch_warp_lu(int ld, int Nmax, const int32_t* Nvec, double* K, int32_t* piv, int32_t* info, GfWork w, int nwork,
                   cudaStream_t s) {
    const int smem = 8 * NMAX * 32 * (int)sizeof(double);
    if (smem > 48 * 1024) cudaFuncSetAttribute(lu_warp_kernel<NMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    lu_warp_kernel<NMAX><<<(nwork + 7) / 8, 256, smem, s>>>(ld, Nvec, Nmax, K, piv, info, w, nwork);
    return gf_launch_status();
}

// ------------------------------------------------------------------------------------------------
template <int NB>
__global__ void __launch_bounds__(256) lu_panel_kernel(int ld, const int32_t* __restrict__ Nvec, int Nfixed,
                                                       double* __restrict__ K, int32_t* __restrict__ piv,
                                                       int32_t* __restrict__ info, GfWork work, int jstart,
                                                       int one_column, int nwork, int trsm_end) {
    // one_column = 0: the whole factorisation (FMA trailing update in this kernel);
    // one_column = 1: block column jstart only -- panel, interchanges, U12 (for the storage rows below trsm_end, < 0:
    // all; the rows beyond get theirs from lu_trsm_kernel after the delayed update); the trailing update is separate
    // The CTAs stride over the work list, so that a (mostly) empty list -- the LU fallback of the LDL' path --
    // costs a small grid instead of `nwork` CTAs that exit at once.
    for (int wi = blockIdx.x; wi < nwork; wi += gridDim.x) {
    const int b = gf_instance(work, wi);
    if (b < 0) return;
    const int N = Nvec != nullptr ? Nvec[b] : Nfixed;
    if (jstart >= N) continue;
    extern __shared__ double P[];              // NB * rows doubles (panel), then 8*NB doubles (U strip)
    double* Us = P + (size_t)NB * ((Nfixed - jstart) | 1);  // Nfixed = batch-wide Nmax (launch_panel)
    __shared__ MaxLoc scratch[32];
    __shared__ int32_t spiv[NB];
    __shared__ int32_t sinfo;
    double* Kb = K + (size_t)b * ld * ld;
    int32_t* pb = piv + (size_t)b * ld;
    const int T = blockDim.x;
    if (threadIdx.x == 0) sinfo = jstart == 0 ? 0 : info[b];
    __syncthreads();
    const int jend = one_column ? min(N, jstart + NB) : N;
    for (int j0 = jstart; j0 < jend; j0 += NB) {
        const int jb = min(NB, N - j0);
        const int rows = N - j0;
        const int pitch = rows | 1;
        for (int c = 0; c < jb; c++)
            for (int i = threadIdx.x; i < rows; i += T) P[c * pitch + i] = Kb[(size_t)(j0 + c) * ld + j0 + i];
        __syncthreads();
        // ---- pivoted unblocked factorisation of the panel in shared memory
        for (int jj = 0; jj < jb; jj++) {
            MaxLoc best{-1.0, rows};
            for (int i = jj + threadIdx.x; i < rows; i += T) {
                const double a = fabs(P[jj * pitch + i]);
                if (a > best.v) { best.v = a; best.i = i; }
            }
            best = block_maxloc(best, scratch);
            const int p = best.i < rows ? best.i : jj;
            if (threadIdx.x == 0) {
                spiv[jj] = j0 + p;
                pb[j0 + jj] = j0 + p;
                record_pivot(best.v, j0 + jj, &sinfo);
            }
            if (p != jj && threadIdx.x < jb) {
                const double t = P[threadIdx.x * pitch + jj];
                P[threadIdx.x * pitch + jj] = P[threadIdx.x * pitch + p];
                P[threadIdx.x * pitch + p] = t;
            }
            __syncthreads();
            const double pv = P[jj * pitch + jj];
            if (pv != 0.0) {
                for (int i = jj + 1 + threadIdx.x; i < rows; i += T) {
                    const double l = P[jj * pitch + i] / pv;
                    P[jj * pitch + i] = l;
                    for (int c = jj + 1; c < jb; c++) P[c * pitch + i] -= l * P[c * pitch + jj];
                }
            }
            __syncthreads();
        }
        for (int c = 0; c < jb; c++)
            for (int i = threadIdx.x; i < rows; i += T) Kb[(size_t)(j0 + c) * ld + j0 + i] = P[c * pitch + i];
        // ---- row interchanges of M on the storage rows outside the panel, then U12 = L11^{-1} M12
        for (int c = (j0 & ~63) + threadIdx.x; c < N; c += T) {  // not the earlier 64-column super-blocks
            if (c >= j0 && c < j0 + jb) continue;
            double* row = Kb + (size_t)c * ld;
            for (int jj = 0; jj < jb; jj++) {
                const int p = spiv[jj];
                if (p != j0 + jj) {
                    const double t = row[j0 + jj];
                    row[j0 + jj] = row[p];
                    row[p] = t;
                }
            }
            if (c >= j0 + jb && (trsm_end < 0 || c < trsm_end)) {
                double u[NB];
#pragma unroll
                for (int k = 0; k < NB; k++) u[k] = (k < jb) ? row[j0 + k] : 0.0;
#pragma unroll
                for (int k = 1; k < NB; k++) {
                    if (k < jb) {
                        double s = u[k];
#pragma unroll
                        for (int q = 0; q < k; q++) s -= P[q * pitch + k] * u[q];
                        u[k] = s;
                    }
                }
#pragma unroll
                for (int k = 0; k < NB; k++)
                    if (k < jb) row[j0 + k] = u[k];
            }
        }
        __syncthreads();
        // ---- trailing update M22 -= L21 U12, strips of 8 storage rows, 4 x 8 register tile per thread
        const int tr = rows - jb;  // trailing extent
        if (tr > 0 && !one_column) {
            for (int c0 = j0 + jb; c0 < N; c0 += 8) {
                const int nc = min(8, N - c0);
                for (int e = threadIdx.x; e < 8 * NB; e += T) {
                    const int r = e / NB, k = e - r * NB;
                    Us[e] = (r < nc && k < jb) ? Kb[(size_t)(c0 + r) * ld + j0 + k] : 0.0;
                }
                __syncthreads();
                for (int i0 = jb + threadIdx.x; i0 < rows; i0 += 4 * T) {
                    double acc[8][4];
#pragma unroll
                    for (int r = 0; r < 8; r++)
#pragma unroll
                        for (int t = 0; t < 4; t++) acc[r][t] = 0.0;
                    for (int k = 0; k < jb; k++) {
                        double pk[4];
#pragma unroll
                        for (int t = 0; t < 4; t++) {
                            const int i = i0 + t * T;
                            pk[t] = (i < rows) ? P[k * pitch + i] : 0.0;
                        }
#pragma unroll
                        for (int r = 0; r < 8; r++) {
                            const double u = Us[r * NB + k];
#pragma unroll
                            for (int t = 0; t < 4; t++) acc[r][t] = fma(pk[t], u, acc[r][t]);
                        }
                    }
#pragma unroll
                    for (int r = 0; r < 8; r++) {
                        if (r < nc) {
                            double* row = Kb + (size_t)(c0 + r) * ld + j0;
#pragma unroll
                            for (int t = 0; t < 4; t++) {
                                const int i = i0 + t * T;
                                if (i < rows) row[i] -= acc[r][t];
                            }
                        }
                    }
                }
                __syncthreads();
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) info[b] = sinfo;
    __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// Block column j0 of the multi-launch factorisation with the panel resident in REGISTERS: thread t owns the R rows of
// the panel that start at logical positions t, t + T, ... (R x NB doubles), no shared-memory copy of the panel, ONE
// block barrier per pivot column.  Rows never move between threads -- every thread tracks the logical position `pos`
// of its rows (the LAPACK interchange jj <-> p swaps two positions).  Pivot search: warp arg-max by shuffles (ties to
// the smaller position, NaN never wins: block_maxloc's rule); the lane that holds a warp's candidate stages that row
// next to the candidate, so after the barrier every warp reduces the <= 32 candidates again and reads the winner's
// row directly.  Then, thread per storage row c outside the panel: the NB interchanges as ONE gather (src[] = inverse
// of `pos`; values that leave the panel range always come from the panel range, so the entering values are held in
// registers and the leaving ones are moved in batches of independent loads) and U12 = L11^{-1} M12.
template <int NB, int R, int TMAX>
__global__ void __launch_bounds__(TMAX) lu_regpanel_kernel(int ld, const int32_t* __restrict__ Nvec, int Nfixed,
                                                           double* __restrict__ K, int32_t* __restrict__ piv,
                                                           int32_t* __restrict__ info, GfWork work, int j0, int nwork,
                                                           int rowops) {  // rowops: U12 for the storage rows below it (< 0: all)
    constexpr int NW = TMAX / 32;
    constexpr int NONE = 1 << 20;
    __shared__ double wrow[2][NW][NB];
    __shared__ double wval[2][NW];
    __shared__ int wpos[2][NW];
    __shared__ double L11[NB][NB + 1];
    __shared__ int spiv[NB];
    __shared__ int src[R * TMAX];
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5, T = blockDim.x, nw = T >> 5;
#pragma unroll 1
    for (int wi = blockIdx.x; wi < nwork; wi += gridDim.x) {
        const int b = gf_instance(work, wi);
        if (b < 0) return;
        const int N = Nvec != nullptr ? Nvec[b] : Nfixed;
        if (j0 >= N) continue;
        const int rows = N - j0, jb = min(NB, rows);
        double* Kb = K + (size_t)b * ld * ld;
        double a[R][NB];
        int pos[R];
#pragma unroll
        for (int r = 0; r < R; r++) {
            pos[r] = t + r * T;
#pragma unroll
            for (int c = 0; c < NB; c++)
                a[r][c] = (pos[r] < rows && c < jb) ? Kb[(size_t)(j0 + c) * ld + j0 + pos[r]] : 0.0;
        }
        int32_t sinfo = j0 == 0 ? 0 : info[b];
#pragma unroll
        for (int jj = 0; jj < NB; jj++) {
            if (jj < jb) {
                double v = -1.0;
                int idx = NONE, rs = 0;
#pragma unroll
                for (int r = 0; r < R; r++) {
                    const double vr = fabs(a[r][jj]);
                    const bool ok = t + r * T < rows && pos[r] >= jj && vr >= 0.0;  // excluded rows, NaN entries
                    if (ok && (vr > v || (vr == v && pos[r] < idx))) { v = vr; idx = pos[r]; rs = r; }
                }
                double bv;
                int bi;
                warp_pivot_search(v, idx, NONE, bv, bi);
                const int buf = jj & 1;
                if (bi == NONE) {
                    if (lane == 0) { wval[buf][wid] = -1.0; wpos[buf][wid] = NONE; }
                } else if (idx == bi) {  // positions are unique: exactly one lane
                    wval[buf][wid] = bv;
                    wpos[buf][wid] = bi;
#pragma unroll
                    for (int c = jj; c < NB; c++) {
                        double x = a[0][c];
#pragma unroll
                        for (int r = 1; r < R; r++)
                            if (rs == r) x = a[r][c];
                        wrow[buf][wid][c] = x;
                    }
                }
                __syncthreads();
                const double cv = lane < nw ? wval[buf][lane] : -1.0;
                const int ci = lane < nw ? wpos[buf][lane] : NONE;
                warp_pivot_search(cv, ci, NONE, bv, bi);
                const int bw = bi < NONE ? __ffs(__ballot_sync(0xffffffffu, ci == bi)) - 1 : 0;  // the warp that staged it
                const int p = bi < NONE ? bi : jj;
                record_pivot(bv, j0 + jj, &sinfo);
                if (t == 0) spiv[jj] = p;
                const double* pr = wrow[buf][bw];
                const double pv = bi < NONE ? pr[jj] : 0.0;
#pragma unroll
                for (int r = 0; r < R; r++) {
                    if (pos[r] == p) pos[r] = jj;
                    else if (pos[r] == jj) pos[r] = p;
                    if (t + r * T < rows && pos[r] > jj && fabs(pv) >= GF_LU_TINY) {
                        const double l = lu_pivot_div(a[r][jj], pv);
                        a[r][jj] = l;
#pragma unroll
                        for (int c = jj + 1; c < NB; c++) a[r][c] = fma(-l, pr[c], a[r][c]);
                    }
                }
            }
        }
#pragma unroll
        for (int r = 0; r < R; r++) {
            if (t + r * T < rows) {
#pragma unroll
                for (int c = 0; c < NB; c++)
                    if (c < jb) Kb[(size_t)(j0 + c) * ld + j0 + pos[r]] = a[r][c];
                src[pos[r]] = t + r * T;
                if (pos[r] < jb) {
#pragma unroll
                    for (int c = 0; c < NB; c++) L11[pos[r]][c] = a[r][c];
                }
            }
        }
        if (t == 0) info[b] = sinfo;
        __syncthreads();
        if (t < jb) piv[(size_t)b * ld + j0 + t] = j0 + spiv[t];
        // ---- interchanges on the storage rows outside the panel + U12
#pragma unroll 1
        for (int c = (j0 & ~63) + t; c < N; c += T) {  // not the earlier 64-column super-blocks
            if (c >= j0 && c < j0 + jb) continue;
            double* rowp = Kb + (size_t)c * ld + j0;
            // all loads of the row are issued before the first dependent use: the entering values (gather by src[])
            // and the old values of the panel range (contiguous), which is where every leaving value comes from
            double uo[NB];  // indexed by src[p] below: lives in local memory (L1), not in registers
#pragma unroll
            for (int k = 0; k < NB; k++) uo[k] = rowp[k];
            double u[NB];
#pragma unroll
            for (int k = 0; k < NB; k++) u[k] = (k < jb) ? rowp[src[k]] : 0.0;
#pragma unroll 4
            for (int jj = 0; jj < jb; jj++) {
                const int p = spiv[jj];
                if (p >= jb) {
                    const int sp = src[p];
                    if (sp != p) rowp[p] = uo[sp];
                }
            }
            if (c >= j0 + jb && (rowops < 0 || c < rowops)) {
#pragma unroll
                for (int k = 1; k < NB; k++) {
                    double sacc = u[k];
#pragma unroll
                    for (int q = 0; q < k; q++) sacc = fma(-L11[k][q], u[q], sacc);
                    u[k] = sacc;
                }
            }
#pragma unroll
            for (int k = 0; k < NB; k++)
                if (k < jb) rowp[k] = u[k];
        }
        __syncthreads();
    }
}

// Row dot with four loads in flight per lane / column axpy term with eight loads in flight per thread (the
// substitution kernels are HBM-bound: without the explicit batching one load per lane is outstanding).
__device__ __forceinline__ double lu_row_dot(const double* __restrict__ row, const double* v, int lo, int hi, int lane) {
    double acc = 0.0;
    int i = lo + lane;
    for (; i + 96 < hi; i += 128) {
        const double r0 = __ldg(row + i), r1 = __ldg(row + i + 32), r2 = __ldg(row + i + 64), r3 = __ldg(row + i + 96);
        acc += r0 * v[i] + r1 * v[i + 32] + r2 * v[i + 64] + r3 * v[i + 96];
    }
    for (; i < hi; i += 32) acc += __ldg(row + i) * v[i];
    return acc;
}

__device__ __forceinline__ double lu_col_dot(const double* __restrict__ col, int ld, const double* v, int jb) {
    double acc = 0.0;
    int jj = 0;
    for (; jj + 8 <= jb; jj += 8) {
        double h[8];
#pragma unroll
        for (int u = 0; u < 8; u++) h[u] = __ldg(col + (size_t)(jj + u) * ld);
#pragma unroll
        for (int u = 0; u < 8; u++) acc += h[u] * v[jj + u];
    }
    for (; jj < jb; jj++) acc += __ldg(col + (size_t)jj * ld) * v[jj];
    return acc;
}

// N <= 32: substitution with one WARP per system, eight systems per CTA, the factor resident in registers: lane r
// holds storage row r (trans = 0: the dots of U' w = r and L' v = w become column sweeps, one broadcast per unknown)
// or storage column r (trans = 1: L z = P r, U x = z), the right-hand side one entry per lane, the interchanges by
// shuffles.  Same operations as lu_solve_kernel in a different summation order.
__global__ void __launch_bounds__(256) lu_solve_warp_kernel(int ld, const int32_t* __restrict__ Nvec, int Nfixed,
                                                            const double* __restrict__ K,
                                                            const int32_t* __restrict__ piv, double* __restrict__ rhs,
                                                            int ldr, int trans, GfWork work, int nwork) {
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slot = blockIdx.x * 8 + wid;
    if (slot >= nwork) return;
    const int b = gf_instance(work, slot);
    if (b < 0) return;
    const int N = Nvec != nullptr ? Nvec[b] : Nfixed;
    if (N <= 0) return;
    const double* Kb = K + (size_t)b * ld * ld;
    double* rb = rhs + (size_t)b * ldr;
    const bool in = lane < N;
    double k[32];
#pragma unroll
    for (int c = 0; c < 32; c++) {
        // trans = 0: k[c] = K[lane][c] (storage row);  trans = 1: k[c] = K[c][lane] (storage column)
        const size_t o = trans ? (size_t)c * ld + lane : (size_t)lane * ld + c;
        k[c] = (in && c < N) ? __ldg(Kb + o) : (c == lane ? 1.0 : 0.0);
    }
    double s = in ? rb[lane] : 0.0;
    const int pj = in ? piv[(size_t)b * ld + lane] : lane;
    if (trans) {  // z = P r
        for (int j = 0; j < N; j++) {
            const int p = __shfl_sync(0xffffffffu, pj, j);
            const double sj = __shfl_sync(0xffffffffu, s, j), sp = __shfl_sync(0xffffffffu, s, p);
            if (lane == j) s = sp;
            else if (lane == p) s = sj;
        }
    }
    // forward sweep: unknown i is final once the sweeps 0..i-1 are applied; lane r > i subtracts its (r, i) entry
    // trans = 0: U' w = r, (r, i) = K[r][i], diagonal K[i][i];  trans = 1: L z = P r, (r, i) = L(r, i) = K[i][r], unit
    double dg = 1.0;
#pragma unroll
    for (int c = 0; c < 32; c++)
        if (c == lane) dg = k[c];
    dg = 1.0 / dg;  // the sweeps multiply by the reciprocal: no division on the serial chain
#pragma unroll
    for (int i = 0; i < 32; i++) {
        if (i < N) {
            double w = trans ? s : s * dg;
            w = __shfl_sync(0xffffffffu, w, i);
            if (lane == i) s = w;
            else if (lane > i) s -= k[i] * w;
        }
    }
    // backward sweep: trans = 0: L' v = w (unit), (r, i) = K[r][i], i > r;  trans = 1: U x = z, (r, i) = K[i][r], diag
#pragma unroll
    for (int i = 31; i >= 0; i--) {
        if (i < N) {
            double w = trans ? s * dg : s;
            w = __shfl_sync(0xffffffffu, w, i);
            if (lane == i) s = w;
            else if (lane < i) s -= k[i] * w;
        }
    }
    if (!trans) {  // x = P' v: the interchanges in reverse
        for (int j = N - 1; j >= 0; j--) {
            const int p = __shfl_sync(0xffffffffu, pj, j);
            const double sj = __shfl_sync(0xffffffffu, s, j), sp = __shfl_sync(0xffffffffu, s, p);
            if (lane == j) s = sp;
            else if (lane == p) s = sj;
        }
    }
    if (in) rb[lane] = s;
}

// 32 < N <= 64: the same idea with two warps per system (thread t holds storage row / column t, 64 doubles): each
// 32-block is swept inside its warp by shuffles, the other block's 32 unknowns arrive through shared memory -- three
// block barriers per substitution instead of one dependent round trip to HBM per block step.  Interchanges block by
// block (factor layout, see the file header): inside warp 1 by shuffles, across the warps in shared memory.
template <bool DIV>
__device__ __forceinline__ void lu_sweep32(double& s, const double* kk, double dg, int lane, int nb, bool up) {
    // kk[0..31]: this thread's coefficients against the 32 unknowns of its own block
#pragma unroll
    for (int q = 0; q < 32; q++) {
        const int i = up ? 31 - q : q;
        if (i < nb) {
            double w = DIV ? s * dg : s;  // dg = reciprocal of the diagonal entry, formed once outside the chain
            w = __shfl_sync(0xffffffffu, w, i);
            if (lane == i) s = w;
            else if (up ? lane < i : lane > i) s -= kk[i] * w;
        }
    }
}

__global__ void __launch_bounds__(64) lu_solve_rows_kernel(int ld, const int32_t* __restrict__ Nvec, int Nfixed,
                                                           const double* __restrict__ K,
                                                           const int32_t* __restrict__ piv, double* __restrict__ rhs,
                                                           int ldr, int trans, GfWork work) {
    const int b = gf_instance(work, blockIdx.x);
    if (b < 0) return;
    const int N = Nvec != nullptr ? Nvec[b] : Nfixed;
    if (N <= 0) return;
    __shared__ double vs[64];
    __shared__ int ps[64];
    const int t = threadIdx.x, lane = t & 31, h = t >> 5;
    const double* Kb = K + (size_t)b * ld * ld;
    double* rb = rhs + (size_t)b * ldr;
    const bool in = t < N;
    double k[64];
#pragma unroll
    for (int c = 0; c < 64; c++) {
        const size_t o = trans ? (size_t)c * ld + t : (size_t)t * ld + c;
        k[c] = (in && c < N) ? __ldg(Kb + o) : (c == t ? 1.0 : 0.0);
    }
    double dg = 1.0;
#pragma unroll
    for (int c = 0; c < 64; c++)
        if (c == t) dg = k[c];
    dg = 1.0 / dg;  // the sweeps multiply: a division per unknown would sit on the serial chain (~370 cycles each)
    vs[t] = in ? rb[t] : 0.0;
    ps[t] = in ? piv[(size_t)b * ld + t] : t;
    const int n0 = min(N, 32), n1 = max(N - 32, 0);  // unknowns in block 0 / block 1
    __syncthreads();
    double s;
    if (!trans) {
        // ---- U' w = r, forward
        s = vs[t];
        if (h == 0) {
            lu_sweep32<true>(s, k, dg, lane, n0, false);
            vs[t] = s;
        }
        __syncthreads();
        if (h == 1) {
#pragma unroll
            for (int i = 0; i < 32; i++) s -= k[i] * vs[i];
            lu_sweep32<true>(s, k + 32, dg, lane, n1, false);
            // ---- L' v = w, backward: block 1, then its interchanges in reverse (all inside this warp)
            lu_sweep32<false>(s, k + 32, 1.0, lane, n1, true);
            for (int jj = n1 - 1; jj >= 0; jj--) {
                const int p = ps[32 + jj] - 32;
                const double sj = __shfl_sync(0xffffffffu, s, jj), sp = __shfl_sync(0xffffffffu, s, p);
                if (lane == jj) s = sp;
                else if (lane == p) s = sj;
            }
            vs[t] = s;
        }
        __syncthreads();
        if (h == 0) {
#pragma unroll
            for (int i = 32; i < 64; i++) s -= k[i] * vs[i];
            lu_sweep32<false>(s, k, 1.0, lane, n0, true);
            vs[t] = s;
            __syncwarp();
            if (lane == 0) {  // block 0's interchanges in reverse: they reach into block 1
                for (int jj = n0 - 1; jj >= 0; jj--) {
                    const int p = ps[jj];
                    if (p != jj) { const double tmp = vs[jj]; vs[jj] = vs[p]; vs[p] = tmp; }
                }
            }
        }
        __syncthreads();
        s = vs[t];
    } else {
        // ---- z = P r block by block, L z = P r forward
        if (t == 0) {
            for (int jj = 0; jj < n0; jj++) {
                const int p = ps[jj];
                if (p != jj) { const double tmp = vs[jj]; vs[jj] = vs[p]; vs[p] = tmp; }
            }
        }
        __syncthreads();
        s = vs[t];
        __syncthreads();
        if (h == 0) {
            lu_sweep32<false>(s, k, 1.0, lane, n0, false);
            vs[t] = s;
        }
        __syncthreads();
        if (h == 1) {
#pragma unroll
            for (int i = 0; i < 32; i++) s -= k[i] * vs[i];
            for (int jj = 0; jj < n1; jj++) {  // block 1's interchanges (inside this warp)
                const int p = ps[32 + jj] - 32;
                const double sj = __shfl_sync(0xffffffffu, s, jj), sp = __shfl_sync(0xffffffffu, s, p);
                if (lane == jj) s = sp;
                else if (lane == p) s = sj;
            }
            lu_sweep32<false>(s, k + 32, 1.0, lane, n1, false);
            // ---- U x = z, backward
            lu_sweep32<true>(s, k + 32, dg, lane, n1, true);
            vs[t] = s;
        }
        __syncthreads();
        if (h == 0) {
#pragma unroll
            for (int i = 32; i < 64; i++) s -= k[i] * vs[i];
            lu_sweep32<true>(s, k, dg, lane, n0, true);
        }
    }
    if (in) rb[t] = s;
}

// The interchanges of one 32-column block applied to the vector (warp 0; v in shared memory): forward order for
// L z = P r, reverse order for x = P' v.
__device__ __forceinline__ void lu_block_swaps(double* v, const int32_t* pb, int j0, int jb, int lane, bool reverse) {
    const int pj = lane < jb ? pb[j0 + lane] : 0;
    for (int s = 0; s < jb; s++) {
        const int jj = reverse ? jb - 1 - s : s;
        const int p = __shfl_sync(0xffffffffu, pj, jj);
        if (lane == 0 && p != j0 + jj) {
            const double t = v[j0 + jj];
            v[j0 + jj] = v[p];
            v[p] = t;
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// Blocked substitution on the packed LU factors; rhs[b] (length >= N) is overwritten by the solution.
// The interchanges are applied block by block (32 columns), matching the factor layout (see the file header).
template <bool RCP>  // RCP: per-lane reciprocal of the diagonal outside the 32-step chains (faster up to N ~ 1000)
__global__ void __launch_bounds__(256) lu_solve_kernel(int ld, const int32_t* __restrict__ Nvec, int Nfixed,
                                                       const double* __restrict__ K, const int32_t* __restrict__ piv,
                                                       double* __restrict__ rhs, int ldr, int trans, GfWork work) {
    const int b = gf_instance(work, blockIdx.x);
    if (b < 0) return;
    const int N = Nvec != nullptr ? Nvec[b] : Nfixed;
    if (N <= 0) return;
    extern __shared__ double v[];        // Nfixed + 1 doubles, then Nfixed pivots
    __shared__ double Tb[32][33];
    __shared__ double part[32];
    const double* Kb = K + (size_t)b * ld * ld;
    int32_t* pb = reinterpret_cast<int32_t*>(v + Nfixed + 1);
    double* rb = rhs + (size_t)b * ldr;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        v[i] = rb[i];
        pb[i] = piv[(size_t)b * ld + i];
    }
    __syncthreads();
    if (!trans) {
        // forward: U' w = r  (storage row j, entries i <= j)
        for (int j0 = 0; j0 < N; j0 += 32) {
            const int jb = min(32, N - j0);
            for (int e = threadIdx.x; e < 32 * 32; e += blockDim.x) {
                const int jj = e >> 5, ii = e & 31;
                Tb[jj][ii] = (jj < jb && ii <= jj) ? Kb[(size_t)(j0 + jj) * ld + j0 + ii] : 0.0;
            }
            for (int jj = wid; jj < jb; jj += nw) {
                const double* row = Kb + (size_t)(j0 + jj) * ld;
                double acc = 0.0;
                acc = lu_row_dot(row, v, 0, j0, lane);
                acc = warp_sum(acc);
                if (lane == 0) part[jj] = acc;
            }
            __syncthreads();
            if (wid == 0) {
                double s = (lane < jb) ? v[j0 + lane] - part[lane] : 0.0;
                const double rd = (RCP && lane < jb) ? 1.0 / Tb[lane][lane] : 0.0;  // one division per lane, off the chain
                for (int ii = 0; ii < jb; ii++) {
                    double w = RCP ? s * rd : s / Tb[ii][ii];
                    w = __shfl_sync(0xffffffffu, w, ii);
                    if (lane == ii) s = w;
                    else if (lane > ii && lane < jb) s -= Tb[lane][ii] * w;
                }
                if (lane < jb) v[j0 + lane] = s;
            }
            __syncthreads();
        }
        // backward: L' x = w (unit; storage row j, entries i > j)
        const int nblk = (N + 31) / 32;
        for (int kb = nblk - 1; kb >= 0; kb--) {
            const int j0 = kb * 32, jb = min(32, N - j0);
            for (int e = threadIdx.x; e < 32 * 32; e += blockDim.x) {
                const int jj = e >> 5, ii = e & 31;
                Tb[jj][ii] = (jj < jb && ii > jj && ii < jb) ? Kb[(size_t)(j0 + jj) * ld + j0 + ii] : 0.0;
            }
            for (int jj = wid; jj < jb; jj += nw) {
                const double* row = Kb + (size_t)(j0 + jj) * ld;
                double acc = 0.0;
                acc = lu_row_dot(row, v, j0 + jb, N, lane);
                acc = warp_sum(acc);
                if (lane == 0) part[jj] = acc;
            }
            __syncthreads();
            if (wid == 0) {
                double s = (lane < jb) ? v[j0 + lane] - part[lane] : 0.0;
                for (int ii = jb - 1; ii >= 0; ii--) {
                    const double w = __shfl_sync(0xffffffffu, s, ii);
                    if (lane < ii) s -= Tb[lane][ii] * w;
                }
                if (lane < jb) v[j0 + lane] = s;
                __syncwarp();
                // the interchanges of a 64-column super-block are applied to the L rows of both of its 32-blocks, so
                // they are undone together, after its first block: second block's in reverse, then the first's
                if ((kb & 1) == 0) {
                    if (j0 + 32 < N) lu_block_swaps(v, pb, j0 + 32, min(32, N - j0 - 32), lane, true);
                    lu_block_swaps(v, pb, j0, jb, lane, true);
                }
            }
            __syncthreads();
        }
    } else {
        // forward: L z = P r  (axpy form; L(i,j) = K[j*ld + i], i > j)
        for (int j0 = 0; j0 < N; j0 += 32) {
            const int jb = min(32, N - j0);
            for (int e = threadIdx.x; e < 32 * 32; e += blockDim.x) {
                const int jj = e >> 5, ii = e & 31;
                Tb[jj][ii] = (jj < jb && ii > jj && ii < jb) ? Kb[(size_t)(j0 + jj) * ld + j0 + ii] : 0.0;
            }
            __syncthreads();
            if (wid == 0) {
                if (((j0 >> 5) & 1) == 0) {  // both 32-blocks' interchanges of the 64-column super-block, up front
                    lu_block_swaps(v, pb, j0, jb, lane, false);
                    if (j0 + 32 < N) lu_block_swaps(v, pb, j0 + 32, min(32, N - j0 - 32), lane, false);
                }
                double s = (lane < jb) ? v[j0 + lane] : 0.0;
                for (int jj = 0; jj < jb; jj++) {
                    const double z = __shfl_sync(0xffffffffu, s, jj);
                    if (lane > jj) s -= Tb[jj][lane] * z;
                }
                if (lane < jb) v[j0 + lane] = s;
            }
            __syncthreads();
            for (int i = j0 + jb + threadIdx.x; i < N; i += blockDim.x) {
                double acc = 0.0;
                acc = lu_col_dot(Kb + (size_t)j0 * ld + i, ld, v + j0, jb);
                v[i] -= acc;
            }
            __syncthreads();
        }
        // backward: U x = z  (U(i,j) = K[j*ld + i], i <= j)
        const int nblk = (N + 31) / 32;
        for (int kb = nblk - 1; kb >= 0; kb--) {
            const int j0 = kb * 32, jb = min(32, N - j0);
            for (int e = threadIdx.x; e < 32 * 32; e += blockDim.x) {
                const int jj = e >> 5, ii = e & 31;
                Tb[jj][ii] = (jj < jb && ii <= jj) ? Kb[(size_t)(j0 + jj) * ld + j0 + ii] : 0.0;
            }
            __syncthreads();
            if (wid == 0) {
                double s = (lane < jb) ? v[j0 + lane] : 0.0;
                const double rd = (RCP && lane < jb) ? 1.0 / Tb[lane][lane] : 0.0;
                for (int jj = jb - 1; jj >= 0; jj--) {
                    double xj = RCP ? s * rd : s / Tb[jj][jj];
                    xj = __shfl_sync(0xffffffffu, xj, jj);
                    if (lane == jj) s = xj;
                    else if (lane < jj) s -= Tb[jj][lane] * xj;
                }
                if (lane < jb) v[j0 + lane] = s;
            }
            __syncthreads();
            for (int i = threadIdx.x; i < j0; i += blockDim.x) {
                double acc = 0.0;
                acc = lu_col_dot(Kb + (size_t)j0 * ld + i, ld, v + j0, jb);
                v[i] -= acc;
            }
            __syncthreads();
        }
    }
    for (int i = threadIdx.x; i < N; i += blockDim.x) rb[i] = v[i];
}

// ------------------------------------------------------------------------------------------------
// Right-looking trailing update on the FP64 tensor pipe: M22 -= L21 U12 for block column j0 (width NB).  In storage
// terms (storage row c = column c of M): C[c][i] -= sum_k U[c][k] L[k][i] with U[c][k] = K[c][j0 + k] (row-major, the
// A operand) and L[k][i] = K[j0 + k][i] (the B operand).  One CTA per 64 (c) x 128 (i) tile, 2 x 4 warps of 32 x 32.
__device__ __forceinline__ double lu_dneg(double x) {
    return __longlong_as_double(__double_as_longlong(x) ^ (long long)0x8000000000000000ULL);
}

template <int NB, int TN>
__global__ void __launch_bounds__(TN * 2, TN == 64 ? 4 : 2) lu_update_kernel(int ld, const int32_t* __restrict__ Nvec,
                                                                              int Nfixed, int j0,
                                                                              double* __restrict__ K, GfWork work,
                                                                              int nwork) {
    // TN = 64: four resident CTAs of four warps per SM instead of two of eight -- the kernel is a single-stage
    // load / DMMA / store sequence per tile, so independent CTAs are what overlaps one tile's loads with another's math
    constexpr int T = TN * 2;
    for (int wi = blockIdx.z; wi < nwork; wi += gridDim.z) {
    const int b = gf_instance(work, wi);
    if (b < 0) return;
    const int N = Nvec != nullptr ? Nvec[b] : Nfixed;
    const int t0 = j0 + NB;
    const int c0 = t0 + blockIdx.y * 64, i0 = t0 + blockIdx.x * TN;
    if (c0 >= N || i0 >= N) continue;
    constexpr int AP = NB + 4;  // pitch of the A tile (== 4 or 12 mod 16: conflict-free fragment loads)
    constexpr int BP = TN + 4;  // pitch of the B tile
    extern __shared__ double usm[];
    double* As = usm;             // 64 x AP
    double* Bs = usm + 64 * AP;   // NB x BP
    double* Kb = K + (size_t)b * ld * ld;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int wm = wid / (TN / 32), wn = wid % (TN / 32), g = lane >> 2, q = lane & 3;
    double acc[4][4][2];
#pragma unroll
    for (int mi = 0; mi < 4; mi++) {
        const int c = c0 + wm * 32 + mi * 8 + g;
#pragma unroll
        for (int ni = 0; ni < 4; ni++) {
            const int i = i0 + wn * 32 + ni * 8 + 2 * q;
            const double* cp = Kb + (size_t)c * ld + i;
            acc[mi][ni][0] = (c < N && i < N) ? cp[0] : 0.0;
            acc[mi][ni][1] = (c < N && i + 1 < N) ? cp[1] : 0.0;
        }
    }
    for (int e = tid; e < 64 * NB; e += T) {
        const int c = e / NB, k = e - c * NB;
        As[c * AP + k] = (c0 + c < N) ? Kb[(size_t)(c0 + c) * ld + j0 + k] : 0.0;
    }
    for (int e = tid; e < NB * TN; e += T) {
        const int k = e / TN, i = e % TN;
        Bs[k * BP + i] = (i0 + i < N) ? Kb[(size_t)(j0 + k) * ld + i0 + i] : 0.0;
    }
    __syncthreads();
    const double* as = As + (wm * 32 + g) * AP + q;
    const double* bs = Bs + q * BP + wn * 32 + g;
#pragma unroll
    for (int kk = 0; kk < NB; kk += 4) {
        double a[4], bf[4];
#pragma unroll
        for (int mi = 0; mi < 4; mi++) a[mi] = lu_dneg(as[mi * 8 * AP + kk]);
#pragma unroll
        for (int ni = 0; ni < 4; ni++) bf[ni] = bs[kk * BP + ni * 8];
#pragma unroll
        for (int mi = 0; mi < 4; mi++)
#pragma unroll
            for (int ni = 0; ni < 4; ni++) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], bf[ni]);
    }
#pragma unroll
    for (int mi = 0; mi < 4; mi++) {
        const int c = c0 + wm * 32 + mi * 8 + g;
#pragma unroll
        for (int ni = 0; ni < 4; ni++) {
            const int i = i0 + wn * 32 + ni * 8 + 2 * q;
            double* cp = Kb + (size_t)c * ld + i;
            if (c < N && i < N) cp[0] = acc[mi][ni][0];
            if (c < N && i + 1 < N) cp[1] = acc[mi][ni][1];
        }
    }
    __syncthreads();  // the shared tiles are reused by the next work item
    }
}

// Generalised trailing update for the delayed (two-level) scheme: C[c][i] -= sum_{k < DEPTH} U[c][k0 + k] L[k0 + k][i]
// for storage rows c in [c_lo, c_hi) and positions i in [i_lo, i_hi) (a bound < 0 means "to the end": N).  Narrow panels
// (8 / 16 / 32 columns) only update the rest of their 64-column super-block -- its remaining storage rows and, for the
// rows beyond, its remaining positions -- with DEPTH = panel width; the bulk of the trailing matrix is touched ONCE per
// super-block with DEPTH = 64: half (32-wide panels) to an eighth (8-wide) of the HBM traffic of updating it after every
// panel, which is what bounded the right-looking update (8 flop per byte instead of 4 ... 1).
template <int DEPTH>
__global__ void __launch_bounds__(128, DEPTH == 64 ? 3 : 4) lu_update_region_kernel(
    int ld, const int32_t* __restrict__ Nvec, int Nfixed, int k0, int depth, int c_lo, int c_hi, int i_lo, int i_hi,
    double* __restrict__ K, GfWork work, int nwork) {
    // depth <= DEPTH (a multiple of 4): the shared tiles are sized for DEPTH
    constexpr int T = 128, TN = 64;
    constexpr int AP = DEPTH + 4;  // pitch of the A tile (== 4 or 12 mod 16: conflict-free fragment loads)
    constexpr int BP = TN + 4;     // pitch of the B tile
    extern __shared__ double usm[];
    double* As = usm;             // 64 x AP
    double* Bs = usm + 64 * AP;   // DEPTH x BP
    for (int wi = blockIdx.z; wi < nwork; wi += gridDim.z) {
        const int b = gf_instance(work, wi);
        if (b < 0) return;
        const int N = Nvec != nullptr ? Nvec[b] : Nfixed;
        const int ce = c_hi < 0 ? N : min(c_hi, N), ie = i_hi < 0 ? N : min(i_hi, N);
        const int c0 = c_lo + blockIdx.y * 64, i0 = i_lo + blockIdx.x * TN;
        if (c0 >= ce || i0 >= ie) continue;
        double* Kb = K + (size_t)b * ld * ld;
        const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
        const int wm = wid >> 1, wn = wid & 1, g = lane >> 2, q = lane & 3;
        double acc[4][4][2];
#pragma unroll
        for (int mi = 0; mi < 4; mi++) {
            const int c = c0 + wm * 32 + mi * 8 + g;
#pragma unroll
            for (int ni = 0; ni < 4; ni++) {
                const int i = i0 + wn * 32 + ni * 8 + 2 * q;
                const double* cp = Kb + (size_t)c * ld + i;
                acc[mi][ni][0] = (c < ce && i < ie) ? cp[0] : 0.0;
                acc[mi][ni][1] = (c < ce && i + 1 < ie) ? cp[1] : 0.0;
            }
        }
        for (int e = tid; e < 64 * depth; e += T) {
            const int c = e / depth, k = e - c * depth;
            As[c * AP + k] = (c0 + c < ce) ? Kb[(size_t)(c0 + c) * ld + k0 + k] : 0.0;
        }
        for (int e = tid; e < depth * TN; e += T) {
            const int k = e / TN, i = e % TN;
            Bs[k * BP + i] = (i0 + i < ie) ? Kb[(size_t)(k0 + k) * ld + i0 + i] : 0.0;
        }
        __syncthreads();
        const double* as = As + (wm * 32 + g) * AP + q;
        const double* bs = Bs + q * BP + wn * 32 + g;
#pragma unroll 4
        for (int kk = 0; kk < depth; kk += 4) {
            double a[4], bf[4];
#pragma unroll
            for (int mi = 0; mi < 4; mi++) a[mi] = lu_dneg(as[mi * 8 * AP + kk]);
#pragma unroll
            for (int ni = 0; ni < 4; ni++) bf[ni] = bs[kk * BP + ni * 8];
#pragma unroll
            for (int mi = 0; mi < 4; mi++)
#pragma unroll
                for (int ni = 0; ni < 4; ni++) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], bf[ni]);
        }
#pragma unroll
        for (int mi = 0; mi < 4; mi++) {
            const int c = c0 + wm * 32 + mi * 8 + g;
#pragma unroll
            for (int ni = 0; ni < 4; ni++) {
                const int i = i0 + wn * 32 + ni * 8 + 2 * q;
                double* cp = Kb + (size_t)c * ld + i;
                if (c < ce && i < ie) cp[0] = acc[mi][ni][0];
                if (c < ce && i + 1 < ie) cp[1] = acc[mi][ni][1];
            }
        }
        __syncthreads();  // the shared tiles are reused by the next work item
    }
}

// Pipelined variant of the region update: persistent CTAs walk the (matrix, tile) list; while the DMMAs of one 64 x 64
// tile run, the operand tiles of the NEXT one stream into the other shared-memory stage (8-byte cp.async with zero fill
// at the ragged edges -- rows of K are only 8-byte aligned when ld is odd) and its C values into registers, so every CTA
// keeps one tile's worth of loads (~100 KB) in flight behind its math instead of alternating load / compute / store.
__device__ __forceinline__ void cp_async8_zfill(void* smem, const void* gmem, bool valid) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    const int n = valid ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(sa), "l"(gmem), "r"(n));
}

template <int DEPTH, bool VEC>
__global__ void __launch_bounds__(256, 1) lu_update_pipe_kernel(
    int ld, const int32_t* __restrict__ Nvec, int Nfixed, int k0, int c_lo, int c_hi, int i_lo, int i_hi, int ntc,
    int nti, double* __restrict__ K, GfWork work, int nwork) {
    // depth == DEPTH.  VEC: ld, k0, i_lo even and the matrices 16-byte aligned -> 16-byte cp.async, else 8-byte.
    // Every CTA owns a contiguous range of the (matrix, tile row, tile column) list and walks it incrementally.
    constexpr int T = 256;
    constexpr int AP = DEPTH + 4, BP = 64 + 4;
    constexpr int STAGE = 64 * AP + DEPTH * BP;
    extern __shared__ double usm[];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int wm = wid >> 2, wn = wid & 3, g = lane >> 2, q = lane & 3;  // 2 x 4 warps of 32 x 16
    const int tpm = ntc * nti;
    const long total = (long)nwork * tpm;
    const long per = (total + gridDim.x - 1) / gridDim.x;
    const long t_begin = (long)blockIdx.x * per, t_end = t_begin + per < total ? t_begin + per : total;
    if (t_begin >= t_end) return;

    struct Tile { double* Kb; int c0, i0, ce, ie; };
    struct Cursor { long t; int wi, tc, ti, b, N; };
    Cursor cu;
    cu.t = t_begin;
    cu.wi = (int)(t_begin / tpm);
    {
        const int r = (int)(t_begin - (long)cu.wi * tpm);
        cu.tc = r / nti;
        cu.ti = r - cu.tc * nti;
    }
    cu.b = -2;
    // advance the cursor to the next tile inside its matrix's range (or past the end); returns false at the end
    auto settle = [&](Cursor& c, Tile& tl) -> bool {
        while (c.t < t_end) {
            if (c.b == -2) {
                c.b = gf_instance(work, c.wi);
                c.N = c.b >= 0 ? (Nvec != nullptr ? Nvec[c.b] : Nfixed) : 0;
            }
            if (c.b >= 0) {
                const int ce = c_hi < 0 ? c.N : min(c_hi, c.N), ie = i_hi < 0 ? c.N : min(i_hi, c.N);
                const int c0 = c_lo + c.tc * 64, i0 = i_lo + c.ti * 64;
                if (c0 < ce && i0 < ie) {
                    tl.Kb = K + (size_t)c.b * ld * ld;
                    tl.c0 = c0; tl.i0 = i0; tl.ce = ce; tl.ie = ie;
                    return true;
                }
            }
            // skip: next tile
            ++c.t;
            if (++c.ti == nti) { c.ti = 0; if (++c.tc == ntc) { c.tc = 0; ++c.wi; c.b = -2; } }
        }
        return false;
    };
    auto step = [&](Cursor& c) {
        ++c.t;
        if (++c.ti == nti) { c.ti = 0; if (++c.tc == ntc) { c.tc = 0; ++c.wi; c.b = -2; } }
    };
    auto fetch_operands = [&](const Tile& tl, int stage) {
        double* As = usm + stage * STAGE;
        double* Bs = As + 64 * AP;
        if (VEC) {
            // A: 64 rows x DEPTH doubles = DEPTH/2 16-byte pieces per row
#pragma unroll
            for (int e = tid; e < 64 * (DEPTH / 2); e += T) {
                const int c = e / (DEPTH / 2), k = (e % (DEPTH / 2)) * 2;
                const bool v = tl.c0 + c < tl.ce;
                const double* src = tl.Kb + (size_t)(v ? tl.c0 + c : tl.c0) * ld + k0 + k;
                const unsigned sa = (unsigned)__cvta_generic_to_shared(As + c * AP + k);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(sa), "l"(src), "r"(v ? 16 : 0));
            }
#pragma unroll
            for (int e = tid; e < DEPTH * 32; e += T) {
                const int k = e >> 5, i = (e & 31) * 2;
                const int rem = tl.ie - (tl.i0 + i);           // doubles of this piece inside the range
                const int nb = rem >= 2 ? 16 : (rem == 1 ? 8 : 0);
                const double* src = tl.Kb + (size_t)(k0 + k) * ld + (nb ? tl.i0 + i : tl.i0);
                const unsigned sa = (unsigned)__cvta_generic_to_shared(Bs + k * BP + i);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(sa), "l"(src), "r"(nb));
            }
        } else {
#pragma unroll 4
            for (int e = tid; e < 64 * DEPTH; e += T) {
                const int c = e / DEPTH, k = e % DEPTH;
                const bool v = tl.c0 + c < tl.ce;
                cp_async8_zfill(As + c * AP + k, tl.Kb + (size_t)(v ? tl.c0 + c : tl.c0) * ld + k0 + k, v);
            }
#pragma unroll 4
            for (int e = tid; e < DEPTH * 64; e += T) {
                const int k = e >> 6, i = e & 63;
                const bool v = tl.i0 + i < tl.ie;
                cp_async8_zfill(Bs + k * BP + i, tl.Kb + (size_t)(k0 + k) * ld + (v ? tl.i0 + i : tl.i0), v);
            }
        }
    };
    auto fetch_c = [&](const Tile& tl, double (&cv)[4][2][2]) {
#pragma unroll
        for (int mi = 0; mi < 4; mi++) {
            const int c = tl.c0 + wm * 32 + mi * 8 + g;
#pragma unroll
            for (int ni = 0; ni < 2; ni++) {
                const int i = tl.i0 + wn * 16 + ni * 8 + 2 * q;
                const double* cp = tl.Kb + (size_t)c * ld + i;
                if (VEC && c < tl.ce && i + 1 < tl.ie) {
                    const double2 v = *reinterpret_cast<const double2*>(cp);
                    cv[mi][ni][0] = v.x;
                    cv[mi][ni][1] = v.y;
                } else {
                    cv[mi][ni][0] = (c < tl.ce && i < tl.ie) ? cp[0] : 0.0;
                    cv[mi][ni][1] = (c < tl.ce && i + 1 < tl.ie) ? cp[1] : 0.0;
                }
            }
        }
    };

    Tile cur, nx;
    if (!settle(cu, cur)) return;
    double acc[4][2][2], nxt[4][2][2];
    int stage = 0;
    fetch_operands(cur, 0);
    cp_async_commit();
    fetch_c(cur, acc);
    while (true) {
        step(cu);
        const bool more = settle(cu, nx);
        cp_async_wait<0>();
        __syncthreads();  // stage `stage` has landed; everybody is done with stage ^ 1 (previous tile's math)
        if (more) {
            fetch_operands(nx, stage ^ 1);
            fetch_c(nx, nxt);
        }
        cp_async_commit();
        const double* As = usm + stage * STAGE;
        const double* Bs = As + 64 * AP;
        const double* as = As + (wm * 32 + g) * AP + q;
        const double* bs = Bs + q * BP + wn * 16 + g;
#pragma unroll
        for (int kk = 0; kk < DEPTH; kk += 4) {
            double a[4], bf[2];
#pragma unroll
            for (int mi = 0; mi < 4; mi++) a[mi] = lu_dneg(as[mi * 8 * AP + kk]);
#pragma unroll
            for (int ni = 0; ni < 2; ni++) bf[ni] = bs[kk * BP + ni * 8];
#pragma unroll
            for (int mi = 0; mi < 4; mi++)
#pragma unroll
                for (int ni = 0; ni < 2; ni++) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], bf[ni]);
        }
#pragma unroll
        for (int mi = 0; mi < 4; mi++) {
            const int c = cur.c0 + wm * 32 + mi * 8 + g;
#pragma unroll
            for (int ni = 0; ni < 2; ni++) {
                const int i = cur.i0 + wn * 16 + ni * 8 + 2 * q;
                double* cp = cur.Kb + (size_t)c * ld + i;
                if (VEC && c < cur.ce && i + 1 < cur.ie) {
                    *reinterpret_cast<double2*>(cp) = make_double2(acc[mi][ni][0], acc[mi][ni][1]);
                } else {
                    if (c < cur.ce && i < cur.ie) cp[0] = acc[mi][ni][0];
                    if (c < cur.ce && i + 1 < cur.ie) cp[1] = acc[mi][ni][1];
                }
            }
        }
        if (!more) break;
#pragma unroll
        for (int mi = 0; mi < 4; mi++)
#pragma unroll
            for (int ni = 0; ni < 2; ni++) { acc[mi][ni][0] = nxt[mi][ni][0]; acc[mi][ni][1] = nxt[mi][ni][1]; }
        cur = nx;
        stage ^= 1;
    }
}

constexpr int LU_GRID_CAP = 256;

template <int DEPTH>
int launch_update_region(int ld, int Nmax, const int32_t* Nvec, double* K, GfWork w, int nwork, cudaStream_t s, int k0,
                         int depth, int c_lo, int c_hi, int i_lo, int i_hi) {
    if (depth <= 0) return GF_OK;
    const int cn = (c_hi < 0 ? Nmax : (c_hi < Nmax ? c_hi : Nmax)) - c_lo, in = (i_hi < 0 ? Nmax : (i_hi < Nmax ? i_hi : Nmax)) - i_lo;
    if (cn <= 0 || in <= 0) return GF_OK;
    // regions with many tiles per matrix (the bulk updates): persistent pipelined CTAs
    static const int pipe_min = getenv("GF_LU_PIPE_MIN") ? atoi(getenv("GF_LU_PIPE_MIN")) : 4;
    const int ntc = (cn + 63) / 64, nti = (in + 63) / 64;
    if (DEPTH == 64 && depth == DEPTH && (long)ntc * nti >= pipe_min && w.count_dev == nullptr) {
        constexpr int PSMEM = 2 * (64 * (DEPTH + 4) + DEPTH * 68) * (int)sizeof(double);
        const long total = (long)nwork * ntc * nti;
        const int grid = total < 148 ? (int)total : 148;
        const bool vec = (ld % 2 == 0) && (k0 % 2 == 0) && (i_lo % 2 == 0) && (((size_t)K) % 16 == 0);
        if (vec) {
            cudaFuncSetAttribute(lu_update_pipe_kernel<DEPTH, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, PSMEM);
            lu_update_pipe_kernel<DEPTH, true><<<grid, 256, PSMEM, s>>>(ld, Nvec, Nmax, k0, c_lo, c_hi, i_lo, i_hi, ntc, nti,
                                                                       K, w, nwork);
        } else {
            cudaFuncSetAttribute(lu_update_pipe_kernel<DEPTH, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, PSMEM);
            lu_update_pipe_kernel<DEPTH, false><<<grid, 256, PSMEM, s>>>(ld, Nvec, Nmax, k0, c_lo, c_hi, i_lo, i_hi, ntc, nti,
                                                                        K, w, nwork);
        }
        return gf_launch_status();
    }
    const int gz = (w.count_dev != nullptr && nwork > LU_GRID_CAP) ? LU_GRID_CAP : nwork;
    dim3 grid((in + 63) / 64, (cn + 63) / 64, gz);
    constexpr int USMEM = (64 * (DEPTH + 4) + DEPTH * 68) * (int)sizeof(double);
    if (USMEM > 48 * 1024)
        cudaFuncSetAttribute(lu_update_region_kernel<DEPTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, USMEM);
    lu_update_region_kernel<DEPTH><<<grid, 128, USMEM, s>>>(ld, Nvec, Nmax, k0, depth, c_lo, c_hi, i_lo, i_hi, K, w, nwork);
    return gf_launch_status();
}

template <int NB>
int launch_panel(int ld, int Nmax, const int32_t* Nvec, int Nfixed, double* K, int32_t* piv, int32_t* info,
                 GfWork w, int nwork, cudaStream_t s, int jstart = 0, int one_column = 0, int trsm_end = -1) {
    const size_t smem = ((size_t)NB * ((Nmax - jstart) | 1) + 8 * NB) * sizeof(double);
    if (smem > 227 * 1024) return GF_ERR_UNSUPPORTED;
    cudaFuncSetAttribute(lu_panel_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    // a device-side count may be far below nwork (fallback lists): bound the grid, the CTAs stride
    const int grid = (w.count_dev != nullptr && nwork > LU_GRID_CAP) ? LU_GRID_CAP : nwork;
    lu_panel_kernel<NB><<<grid, 256, smem, s>>>(ld, Nvec, Nfixed, K, piv, info, w, jstart, one_column, nwork, trsm_end);
    return gf_launch_status();
}

// One block column of the multi-launch factorisation: pivoted panel (one CTA per matrix), then the DMMA update.
// R > 0: register-resident panel, R rows per thread (rows left <= 512 R); R = 0: the shared-memory panel.
// developer timing (GF_LU_TIMING=1): CUDA events around every launch of the multi-launch path, summed per kernel type
struct LuTiming {
    bool on;
    std::vector<cudaEvent_t> ev;   // triples are not needed: events alternate begin / end
    std::vector<int> kind;         // 0 panel, 1 update
};
static LuTiming g_lu_timing = {getenv("GF_LU_TIMING") != nullptr, {}, {}};
static void lu_time_mark(cudaStream_t s, int kind) {
    if (!g_lu_timing.on) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, s);
    g_lu_timing.ev.push_back(e);
    g_lu_timing.kind.push_back(kind);
}
static void lu_time_report(cudaStream_t s) {
    if (!g_lu_timing.on || g_lu_timing.ev.empty()) return;
    cudaStreamSynchronize(s);
    float tot[2] = {0.f, 0.f};
    for (size_t i = 0; i + 1 < g_lu_timing.ev.size(); i += 2) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, g_lu_timing.ev[i], g_lu_timing.ev[i + 1]);
        tot[g_lu_timing.kind[i]] += ms;
    }
    float all = 0.f;
    cudaEventElapsedTime(&all, g_lu_timing.ev.front(), g_lu_timing.ev.back());
    fprintf(stderr, "[gf_lu timing] panel %.3f ms, update %.3f ms, first-to-last %.3f ms, %zu launches\n", tot[0], tot[1], all,
            g_lu_timing.ev.size() / 2);
    for (cudaEvent_t e : g_lu_timing.ev) cudaEventDestroy(e);
    g_lu_timing.ev.clear();
    g_lu_timing.kind.clear();
}

// U12 of block column j0 for the storage rows c >= c_lo (the columns beyond the panel's 64-column super-block), after
// their delayed update: u = L11^{-1} K[c][j0 .. j0 + jb), forward substitution in the order of the panel kernels' own
// U12 step.  Thread per storage row, L11 (unit lower; position k of column q = K[j0 + q][j0 + k]) in shared memory.
template <int NB>
__global__ void __launch_bounds__(256) lu_trsm_kernel(int ld, const int32_t* __restrict__ Nvec, int Nfixed, int j0,
                                                      int c_lo, double* __restrict__ K, GfWork work, int nwork) {
    __shared__ double L11[NB][NB + 1];
    for (int wi = blockIdx.y; wi < nwork; wi += gridDim.y) {
        const int b = gf_instance(work, wi);
        if (b < 0) return;
        const int N = Nvec != nullptr ? Nvec[b] : Nfixed;
        if (c_lo >= N || j0 >= N) continue;
        const int jb = min(NB, N - j0);
        double* Kb = K + (size_t)b * ld * ld;
        for (int e = threadIdx.x; e < NB * NB; e += blockDim.x) {
            const int k = e / NB, q = e - k * NB;
            L11[k][q] = (k < jb && q < k) ? Kb[(size_t)(j0 + q) * ld + j0 + k] : 0.0;
        }
        __syncthreads();
        const int c = c_lo + blockIdx.x * blockDim.x + threadIdx.x;
        if (c < N) {
            double* rowp = Kb + (size_t)c * ld + j0;
            double u[NB];
#pragma unroll
            for (int k = 0; k < NB; k++) u[k] = (k < jb) ? rowp[k] : 0.0;
#pragma unroll
            for (int k = 1; k < NB; k++) {
                double sacc = u[k];
#pragma unroll
                for (int q = 0; q < k; q++) sacc = fma(-L11[k][q], u[q], sacc);
                u[k] = sacc;
            }
#pragma unroll
            for (int k = 0; k < NB; k++)
                if (k < jb) rowp[k] = u[k];
        }
        __syncthreads();
    }
}

// Block column j0 (width NB) of the multi-launch factorisation, delayed-update scheme by 64-column super-blocks:
//   1. pivoted panel (one CTA per matrix): factorisation, interchanges on every storage row of its super-block and
//      beyond, U12 for the later storage rows INSIDE the super-block (they are kept fully up to date, step 4) -- and
//      for all later rows when the panel is the first of its super-block (nothing is pending then);
//   2. storage rows beyond the super-block: the interchanges have just moved positions of the panel's range in from
//      anywhere below, so only now the pending contributions of the super-block's earlier panels (depth j0 - jS) are
//      applied to that range, left-looking;
//   3. ... and their U12 (lu_trsm_kernel);
//   4. the rest of the super-block's own storage rows is updated right away (depth NB, all later positions).
// The bulk -- storage rows and positions beyond the super-block -- is updated once per super-block with depth 64
// (gf_lu_factor), half to an eighth of the HBM traffic of updating it after every panel.
template <int NB, int R>
int launch_column(int ld, int Nmax, const int32_t* Nvec, double* K, int32_t* piv, int32_t* info, GfWork w, int nwork,
                  cudaStream_t s, int j0) {
    // Delayed updates pay from N = 512 on (B200, B = 4096 / 1024: N = 512 53.4 -> 50.0 ms, N = 1024 106 -> 87 ms,
    // N = 2048 1064 -> 551 ms); below, the plain right-looking update is as fast.  GF_LU_DELAY=0 / 1 forces one or the
    // other.
    static const int delay_req = getenv("GF_LU_DELAY") ? atoi(getenv("GF_LU_DELAY")) : -1;
    const bool nodelay = delay_req == 0 || (delay_req < 0 && Nmax < 512);
    const int jS = j0 & ~63, jn = j0 + NB;
    const int send = (jS + 64) < Nmax ? (jS + 64) : Nmax;
    const bool first = (j0 == jS) || nodelay;
    const int trsm_end = first ? -1 : send;
    int rc;
    lu_time_mark(s, 0);
    if constexpr (R > 0) {
        int threads = ((Nmax - j0 + R - 1) / R + 31) & ~31;
        threads = threads < 256 ? 256 : (threads > 512 ? 512 : threads);
        const int grid = (w.count_dev != nullptr && nwork > LU_GRID_CAP) ? LU_GRID_CAP : nwork;
        lu_regpanel_kernel<NB, R, 512><<<grid, threads, 0, s>>>(ld, Nvec, Nmax, K, piv, info, w, j0, nwork, trsm_end);
        rc = gf_launch_status();
    } else {
        rc = launch_panel<NB>(ld, Nmax, Nvec, Nmax, K, piv, info, w, nwork, s, j0, 1, trsm_end);
    }
    lu_time_mark(s, 0);
    if (rc != GF_OK) return rc;
    if (nodelay) {  // the plain right-looking update of the whole trailing matrix (fully unrolled depth-NB kernel)
        const int tr = Nmax - jn;
        if (tr <= 0) return GF_OK;
        lu_time_mark(s, 1);
        const int gz = (w.count_dev != nullptr && nwork > LU_GRID_CAP) ? LU_GRID_CAP : nwork;
        dim3 grid((tr + 63) / 64, (tr + 63) / 64, gz);
        constexpr int USMEM = (64 * (NB + 4) + NB * 68) * (int)sizeof(double);
        lu_update_kernel<NB, 64><<<grid, 128, USMEM, s>>>(ld, Nvec, Nmax, j0, K, w, nwork);
        lu_time_mark(s, 1);
        return gf_launch_status();
    }
    lu_time_mark(s, 1);
    if (!first && send < Nmax) {
        rc = launch_update_region<64>(ld, Nmax, Nvec, K, w, nwork, s, jS, j0 - jS, send, -1, j0, jn < send ? jn : send);
        if (rc != GF_OK) return rc;
        const int rows = Nmax - send;
        const int gy = (w.count_dev != nullptr && nwork > LU_GRID_CAP) ? LU_GRID_CAP : nwork;
        lu_trsm_kernel<NB><<<dim3((rows + 255) / 256, gy), 256, 0, s>>>(ld, Nvec, Nmax, j0, send, K, w, nwork);
        rc = gf_launch_status();
        if (rc != GF_OK) return rc;
    }
    if (jn < send) rc = launch_update_region<NB>(ld, Nmax, Nvec, K, w, nwork, s, j0, NB, jn, send, jn, -1);
    lu_time_mark(s, 1);
    return rc;
}

// Lanes (experiment, off by default): the batch split into parts whose launch sequences run on separate streams, so that
// one part's update could run beside another part's panel.  It cannot: see GF_LU_LANES in gf_lu_factor.
constexpr int LU_MAX_LANES = 8;
constexpr int LU_LANES_MIN = 512;  // matrices from which the split pays
struct LuLanes {
    cudaStream_t s[LU_MAX_LANES];
    cudaEvent_t fork, join[LU_MAX_LANES];
    bool ok;
};
LuLanes* lu_lanes() {
    static std::mutex mu;
    static LuLanes lanes[64];
    static bool made[64] = {false};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    std::lock_guard<std::mutex> lock(mu);
    LuLanes& L = lanes[dev];
    if (!made[dev]) {
        made[dev] = true;
        L.ok = cudaEventCreateWithFlags(&L.fork, cudaEventDisableTiming) == cudaSuccess;
        for (int i = 0; i < LU_MAX_LANES; i++) {
            L.ok = L.ok && cudaStreamCreateWithFlags(&L.s[i], cudaStreamNonBlocking) == cudaSuccess;
            L.ok = L.ok && cudaEventCreateWithFlags(&L.join[i], cudaEventDisableTiming) == cudaSuccess;
        }
    }
    return L.ok ? &L : nullptr;
}

}  // namespace

extern "C" int gf_lu_factor(int B, int ld, int Nmax, const int32_t* Nvec, double* K, int32_t* piv, int32_t* info,
                            const int32_t* work, const int32_t* nwork_dev, int nwork, void* stream) {
    if (B <= 0 || ld <= 0 || Nmax < 0 || Nmax > ld || !K || !piv || !info) return GF_ERR_ARG;
    if (nwork <= 0 || Nmax == 0) return GF_OK;
    cudaStream_t s = (cudaStream_t)stream;
    GfWork w{work, nwork_dev};
    // warp-per-matrix, register-resident
    if (Nmax <= 8) return launch_warp_lu<8>(ld, Nmax, Nvec, K, piv, info, w, nwork, s);
    if (Nmax <= 16) return launch_warp_lu<16>(ld, Nmax, Nvec, K, piv, info, w, nwork, s);
    if (Nmax <= 32) return launch_warp_lu<32>(ld, Nmax, Nvec, K, piv, info, w, nwork, s);
    if (Nmax <= 64) {  // register-resident rows, two warps per matrix
        constexpr int RSMEM = (64 * 64 + 2 * 2 * 64) * (int)sizeof(double);
        cudaFuncSetAttribute(lu_rows_kernel<64, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, RSMEM);
        const int grid = nwork < 148 * 6 ? nwork : 148 * 6;
        lu_rows_kernel<64, 2><<<grid, 64, RSMEM, s>>>(ld, Nvec, Nmax, K, piv, info, w, nwork);
        return gf_launch_status();
    }
    const size_t full = (size_t)Nmax * (Nmax | 1) * sizeof(double);
    if (full <= 100 * 1024) {
        if (full > 48 * 1024) cudaFuncSetAttribute(lu_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)full);
        const int threads = Nmax <= 32 ? 32 : (Nmax <= 64 ? 64 : 128);
        lu_smem_kernel<<<nwork, threads, full, s>>>(ld, Nvec, Nmax, K, piv, info, w);
        return gf_launch_status();
    }
    if (Nmax > 3500 || nwork > 65535) return GF_ERR_UNSUPPORTED;
    // A device-counted work list is the fallback list of the LDL' path: usually empty, at most a few instances.
    // One launch of the whole-factorisation kernel (bounded grid, FMA trailing update) instead of two per block
    // column keeps its cost at a few microseconds when there is nothing to do.
    if (nwork_dev != nullptr && Nmax <= 830) return launch_panel<32>(ld, Nmax, Nvec, Nmax, K, piv, info, w, nwork, s);
    // Multi-launch right-looking factorisation: per block column a pivoted panel kernel (panel resident in shared
    // memory, so its width follows the rows that are left) and the DMMA trailing update.
    // GF_LU_LANES: parts of the batch whose launch sequences interleave on separate streams.  Measured: no gain (N = 512,
    // B = 4096: 53.4 ms for 1, 2, 4 and 8 lanes) -- the panel kernel holds all of an SM's registers, so another part's
    // update cannot co-reside with it; off by default (1), kept for the record
    static const int lanes_req = [] {
        const char* e = getenv("GF_LU_LANES");
        const int v = e == nullptr ? 1 : atoi(e);
        return v < 2 ? 1 : (v > LU_MAX_LANES ? LU_MAX_LANES : v);
    }();
    LuLanes* L = (lanes_req > 1 && nwork >= LU_LANES_MIN && nwork_dev == nullptr) ? lu_lanes() : nullptr;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (L != nullptr && (cudaStreamIsCapturing(s, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone)) L = nullptr;
    static std::mutex issue_mu;  // the lane streams / events are per device: one caller issues at a time (see gf_ldlt.cu)
    std::unique_lock<std::mutex> issue_lock(issue_mu, std::defer_lock);
    const int nlane = L != nullptr ? lanes_req : 1;
    int off[LU_MAX_LANES], cnt[LU_MAX_LANES];
    cudaStream_t st[LU_MAX_LANES];
    for (int i = 0; i < nlane; i++) {
        off[i] = (int)((long)nwork * i / nlane);
        cnt[i] = (int)((long)nwork * (i + 1) / nlane) - off[i];
        st[i] = L != nullptr ? L->s[i] : s;
    }
    if (L != nullptr) {
        issue_lock.lock();
        cudaEventRecord(L->fork, s);
        for (int i = 0; i < nlane; i++) cudaStreamWaitEvent(L->s[i], L->fork, 0);
    }
    int j0 = 0, nb = 8;
    while (j0 < Nmax) {
        const int rows = Nmax - j0;
        // widest panel the rows left allow; a wider panel starts only on its own multiple (panels never straddle a
        // 32-column block: the interchanges are not applied to the blocks left of the current one)
        // (the panel lives in registers up to 512 R rows, in shared memory -- 16 wide up to 1700 rows -- beyond)
        const int want = rows <= 512 ? 32 : (rows <= 1700 ? 16 : 8);
        if (want <= nb || j0 % want == 0) nb = want;
        for (int i = 0; i < nlane; i++) {
            GfWork wl{work, nwork_dev, off[i]};
            int rc;
            if (nb == 32) rc = launch_column<32, 1>(ld, Nmax, Nvec, K, piv, info, wl, cnt[i], st[i], j0);
            else if (nb == 16 && rows <= 1024) rc = launch_column<16, 2>(ld, Nmax, Nvec, K, piv, info, wl, cnt[i], st[i], j0);
            else if (nb == 16) rc = launch_column<16, 0>(ld, Nmax, Nvec, K, piv, info, wl, cnt[i], st[i], j0);
            else if (rows <= 2048) rc = launch_column<8, 4>(ld, Nmax, Nvec, K, piv, info, wl, cnt[i], st[i], j0);
            else rc = launch_column<8, 0>(ld, Nmax, Nvec, K, piv, info, wl, cnt[i], st[i], j0);
            if (rc != GF_OK) return rc;
            const int jn = j0 + nb;
            static const int delay_req = getenv("GF_LU_DELAY") ? atoi(getenv("GF_LU_DELAY")) : -1;
            const bool nodelay = delay_req == 0 || (delay_req < 0 && Nmax < 512);
            if (!nodelay && (jn & 63) == 0 && jn < Nmax) {  // the super-block [jn - 64, jn) is complete: its one depth-64 update
                lu_time_mark(st[i], 1);
                rc = launch_update_region<64>(ld, Nmax, Nvec, K, wl, cnt[i], st[i], jn - 64, 64, jn, -1, jn, -1);
                lu_time_mark(st[i], 1);
                if (rc != GF_OK) return rc;
            }
        }
        j0 += nb;
    }
    if (L != nullptr) {
        for (int i = 0; i < nlane; i++) {
            cudaEventRecord(L->join[i], L->s[i]);
            cudaStreamWaitEvent(s, L->join[i], 0);
        }
    }
    lu_time_report(s);
    return GF_OK;
}

extern "C" int gf_lu_solve(int B, int ld, int Nmax, const int32_t* Nvec, const double* K, const int32_t* piv,
                           double* rhs, int ldr, int trans, const int32_t* work, const int32_t* nwork_dev, int nwork,
                           void* stream) {
    if (B <= 0 || ld <= 0 || Nmax < 0 || Nmax > ld || ldr < Nmax || !K || !piv || !rhs) return GF_ERR_ARG;
    if (nwork <= 0 || Nmax == 0) return GF_OK;
    if (Nmax <= 32) {
        lu_solve_warp_kernel<<<(nwork + 7) / 8, 256, 0, (cudaStream_t)stream>>>(ld, Nvec, Nmax, K, piv, rhs, ldr, trans,
                                                                               GfWork{work, nwork_dev}, nwork);
        return gf_launch_status();
    }
    if (Nmax <= 64) {
        lu_solve_rows_kernel<<<nwork, 64, 0, (cudaStream_t)stream>>>(ld, Nvec, Nmax, K, piv, rhs, ldr, trans,
                                                                    GfWork{work, nwork_dev});
        return gf_launch_status();
    }
    const size_t smem = (size_t)(Nmax + 1) * sizeof(double) + (size_t)Nmax * sizeof(int32_t);
    if (smem > 200 * 1024) return GF_ERR_UNSUPPORTED;
    if (Nmax < 1000) {
        lu_solve_kernel<true><<<nwork, 256, smem, (cudaStream_t)stream>>>(ld, Nvec, Nmax, K, piv, rhs, ldr, trans,
                                                                          GfWork{work, nwork_dev});
    } else {  // measured: from N ~ 1000 the kernel with the divisions in the chains is the faster one (HBM-bound)
        if (smem > 48 * 1024)
            cudaFuncSetAttribute(lu_solve_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        lu_solve_kernel<false><<<nwork, 256, smem, (cudaStream_t)stream>>>(ld, Nvec, Nmax, K, piv, rhs, ldr, trans,
                                                                           GfWork{work, nwork_dev});
    }
    return gf_launch_status();
}
