// Batched dense FP64 LDL' (no pivoting) for symmetric quasi-definite KKT matrices on the FP64 tensor pipe
// (mma.sync m8n8k4 f64 -> SASS DMMA.8x8x4), + substitution (a13, a14, inertia contract).
//
// Replaces pygradflow/linear_solver/lu_solver.py:9-21 on the Symmetric step solver's matrix
// K = [[H_II + lamb I, J_I'], [J_I, -delta I]] (symmetric_step_solver.py:49-77), which is quasi-definite
// whenever H_II + lamb I > 0, so L D L' exists for every symmetric permutation and is stable without
// pivoting; the signs of D give the inertia that ma57_solver.py:76-79 / mumps_solver.py:81-82 /
// ssids_solver.py:22-23 report (num_neg_eigvals).
//
// Layout: K[b] ld x ld row-major, only the lower triangle is read; the matrix is padded with an identity
// block up to Np = roundup(N_b, 64) (gf_kkt_assemble does this) so no kernel has edge tiles inside a
// 64-block.  d is kept on the diagonal of K and, contiguously, in dvec[b].  The upper triangle of K is scratch:
// the strict upper triangle of every 64x64 DIAGONAL block receives inv(L_kk)' for the triangular solves, and the
// off-diagonal block (p, i), p < i, receives the block W'[i,p] = -(L[i,p] D_p) (not transposed), so the update
// product sum_p W'[i,p] L[k,p]' needs no scaling inside the DMMA loop (a DMUL in that loop shares the FP64
// pipe with the DMMAs and costs 8 % of its throughput, tools/dmma_loop_bench.cu).
//
// Left-looking by 64-wide block columns, ONE batched launch per block column k (j0 = 64 k) after the first
// diagonal block; per matrix the launch has
//   chain CTA   the 64 x 128 tile [ A[k+1,k] | A[k+1,k+1] ] minus the contributions of the block columns to
//               its left in one DMMA main loop (the rows of block k+1 are both the A operand and the B operand of
//               the diagonal tile, so it costs no extra loads), the triangular solve of L[k+1,k], the last
//               rank-64 update of the diagonal tile, and then in shared memory C = L D L' and X = inv(L)
//               (blocked 16/32/64) of diagonal block k+1  -- the critical path of the factorisation;
//   panel CTAs  128x64 tiles further below: C = A[i,k] - sum_{p<k} (L[i,p] D_p) L[k,p]'  (DMMA), then the
//               triangular solve as one more DMMA product L[i,k] = (C X') D_k^{-1}.
// Every matrix entry is read and written once, operands stream once per block column
// (~ 8 N^3 / (6*64) bytes per matrix), so the factorisation is bound by the FP64 pipe, not by HBM.
#include "gf_common.cuh"
#include "../../include/gradflow_b200.h"
#include <cstdlib>
#include <cstring>
#include <mutex>

#ifdef GF_LDLT_TRACE
// Developer instrumentation (tools/ldlt_trace.cu): per block column / CTA role, cycles spent per phase.
__device__ unsigned long long g_ldlt_trace[64 * 2 * 16];
#define TRACE_BEGIN(kk, role) long long t_prev_ = clock64(); const int t_slot_ = ((kk) * 2 + (role)) * 16;
#define TRACE_MARK(i)                                                                    \
    do {                                                                                 \
        if (threadIdx.x == 0) {                                                          \
            const long long t_now_ = clock64();                                          \
            atomicAdd(&g_ldlt_trace[t_slot_ + (i)], (unsigned long long)(t_now_ - t_prev_)); \
            t_prev_ = t_now_;                                                            \
        }                                                                                \
    } while (0)
#define TRACE_COUNT() do { if (threadIdx.x == 0) atomicAdd(&g_ldlt_trace[t_slot_ + 15], 1ULL); } while (0)
__device__ unsigned long long g_ldlt_fine[8];
#define FINE_DECL __shared__ long long fine_ts_[80]; int fine_n_ = 0;
#define FINE_MARK() do { if (threadIdx.x == 0 && fine_n_ < 80) fine_ts_[fine_n_] = clock64(); fine_n_++; } while (0)
#define FINE_FLUSH() do { if (threadIdx.x == 0) { for (int i_ = 1; i_ < fine_n_ && i_ < 80; i_++) atomicAdd(&g_ldlt_fine[(i_ - 1) % 5], (unsigned long long)(fine_ts_[i_] - fine_ts_[i_ - 1])); } } while (0)
#else
#define FINE_DECL
#define FINE_MARK()
#define FINE_FLUSH()
#define TRACE_BEGIN(kk, role)
#define TRACE_MARK(i)
#define TRACE_COUNT()
#endif

namespace {

constexpr int NB = 64;
constexpr int EP = NB + 4;   // smem row pitch of 64-wide tiles in the epilogues (same property)

__device__ __forceinline__ int padded_order(const int32_t* Nvec, int Nfixed, int b, int ld) {
    const int N = Nvec != nullptr ? Nvec[b] : Nfixed;
    const int Np = ((N + NB - 1) / NB) * NB;
    return Np < ld ? Np : ld;
}

// Optional on-the-fly assembly of the KKT matrix (the fused entry gf_kkt_ldlt_factor): instead of reading the
// assembled lower triangle from K, the first touch of every tile gathers it from H, J and the index sets, exactly as
// kkt_assemble_kernel would have written it (symmetric_step_solver.py:27-39,49-77) -- this saves writing and
// re-reading K once (9.6 GB each at cfg3).  H == nullptr: K already holds the assembled matrix.
struct KktSrc {
    const double* H;
    const double* J;
    const int32_t* perm;
    const int32_t* nI;
    const double* dt;
    const double* rho;
    int n, m;
};

struct KktView {  // per-instance view of KktSrc
    const double* Hb;
    const double* Jb;
    const int32_t* pb;
    int nI, N, n;
    double lamb, corner;
    bool on;
};

__device__ __forceinline__ KktView kkt_view(const KktSrc& s, int b) {
    KktView v;
    v.on = s.H != nullptr;
    if (!v.on) return v;
    v.Hb = s.H + (size_t)b * s.n * s.n;
    v.Jb = s.J != nullptr ? s.J + (size_t)b * s.m * s.n : nullptr;
    v.pb = s.perm + (size_t)b * s.n;
    v.nI = s.nI[b];
    v.N = v.nI + s.m;
    v.n = s.n;
    v.lamb = 1.0 / s.dt[b];                                   // symmetric_step_solver.py:30
    v.corner = -v.lamb / (1.0 + v.lamb * s.rho[b]);           // :60-62
    return v;
}

// entry (r, c) of the padded, reduced KKT matrix, lower triangle (c > r: never used, 0)
__device__ __forceinline__ double kkt_lower(const KktView& v, int r, int c) {
    if (c > r) return 0.0;
    if (r < v.nI) {
        const double h = __ldg(v.Hb + (size_t)__ldg(v.pb + r) * v.n + __ldg(v.pb + c));
        return r == c ? __dadd_rn(h, v.lamb) : h;             // H + diag(lamb)  :34-36
    }
    if (r < v.N) {
        if (c < v.nI) return __ldg(v.Jb + (size_t)(r - v.nI) * v.n + __ldg(v.pb + c));
        return r == c ? v.corner : 0.0;
    }
    return r == c ? 1.0 : 0.0;                                // identity padding up to the 64-block
}

__device__ __forceinline__ double2 kkt_load2(const KktView& v, const double* __restrict__ Kb, int ld, int r, int c) {
    if (!v.on) return *reinterpret_cast<const double2*>(Kb + (size_t)r * ld + c);
    return make_double2(kkt_lower(v, r, c), kkt_lower(v, r, c + 1));
}

// Sign flip on the integer pipe (the FP64 pipe is the contended resource here).
__device__ __forceinline__ double dneg(double x) {
    return __longlong_as_double(__double_as_longlong(x) ^ (long long)0x8000000000000000ULL);
}

// ------------------------------------------------------------------------------------------------
// Diagonal block: S = L D L' and X = inv(L) for a 64 x 64 tile in shared memory.
//
// A latency problem: 64 dependent pivots.  The FP64 pipe is shared with the DMMAs of the CTA next door and the warp
// scheduler favours a warp that streams DMMAs over one that waits on a dependent result
// (tools/fp64_contention_bench.cu, profiles/r02_fp64_contention.txt: a dependent DFMA / DMMA takes 9 / 27 cycles alone,
// 137 / 280 next to one, 26 000 / 1 000 next to two streaming warps on the sub-partition -- it does not matter which of
// the two instructions the chain is made of).  The scheme keeps the chain short and everything else off it:
//   * 8 panels of 8 columns.  Warp 0 holds the 8 x 8 pivot block P in registers in the DMMA C layout (lane (g, q) owns
//     P[g][2q], P[g][2q+1]).  Column j of P then lives in the lanes with q == j / 2, which are exactly the lanes that supply
//     k-slot j / 2 of an A fragment (rows) and of a B fragment (columns): the rank-1 update of a pivot step is ONE DMMA
//     with operands taken from the owning lanes' registers, no exchange.  Per pivot: broadcast d_j (one shuffle), MUFU
//     reciprocal seed, e = 1 - d r0, sc = e + e^2, l = (w r0)(1 + sc), DMMA.  Y = inv(L_pivot) is updated the same way
//     (its B operand, row j of Y, comes from Z = Y', kept alongside).  ~200 cycles per pivot (round 1: six 64-bit
//     shuffles per pivot and scalar updates, 275);
//   * panel below the pivot block: W = S_panel Y' (2 DMMAs per 8 rows), L = W D^{-1};
//   * trailing update S += (-W) L' by 8 x 8 tiles (2 DMMAs each), the next pivot block first, by warp 0, which
//     keeps it in registers and goes straight on to eliminate it while the other warps finish the update;
//   * X = inv(L) by block rows (DMMA products of 8 x 8 tiles) by warps 1..7 while warp 0 is busy with the next pivot
//     block: only the last block row is left after the loop (the block recursion 8 -> 16 -> 32 -> 64 of round 1 cost
//     7-16 k cycles per diagonal block after the factorisation).
// Measured per diagonal block, alone on the SM (tools/ldlt_trace, profiles/r02_ldlt_trace_{before,after}.txt):
// factor 24.3 k -> 19.5 k cycles, inverse 7.0 k -> 0.9 k.  Tried, not adopted: the row tiles below the pivot block eliminated
// TOGETHER with it by four warps (the same rank-1 DMMAs applied to the row tile, no inv(L_pivot) on the critical path, no
// panel phase, trailing update overlapped with the next elimination): every DMMA a chain warp issues costs it ~35 cycles,
// three per pivot made the elimination 2.15 k cycles per panel and the block 24 k; warp 0 only signalling the second
// barrier of a panel (bar.arrive) instead of waiting at it: slower (27.9 k).
constexpr int DP = NB + 4;                   // pitch of S and X (== 4 mod 16: conflict-free DMMA fragment loads)
constexpr int WNP = 12;                      // pitch of the -W panel (64 x 8)
constexpr int TP = 36;                       // pitch of the 32 x 32 product scratch
constexpr int DG_S = NB * DP;                // S: 64 x 68
constexpr int DG_OPS = NB * DP;              // X = inv(L): 64 x 68
constexpr int DG_MISC = NB * WNP + 32 * TP + NB + 2;  // -W panel, product scratch, 1/d, flags
constexpr int DG_SMEM = (DG_S + DG_OPS + DG_MISC) * (int)sizeof(double);

// Factorise the updated diagonal tile of block k, held in shared memory S[64][DP] (lower triangle valid):
// S = L D L', X = inv(L); writes L / d / inv(L)' to K, d to dvec, pivot diagnostics to info / nneg.
// Xp (DG_OPS doubles) and misc (DG_MISC doubles) are scratch regions disjoint from S.
// MERGED (the small-order kernel, where shared memory decides how many matrices an SM holds): X is not a second tile --
// inv(L)' lives in the strict upper triangle of S itself, as in the global layout (X[r][c], r > c, at S[c][r]; unit
// diagonal and zeros implied), the 8 x 8 product of x_row_tile changes layout by shuffles instead of through a scratch
// tile, the -W panel has pitch 10: S + 706 doubles (40 KB) instead of 2 tiles + 1986 doubles (85 KB).  Same arithmetic
// in the same order: both variants produce the same bits.
constexpr int WNPM = 10;
constexpr int DGM_MISC = NB * WNPM + NB + 2;  // -W panel, 1/d, flags
template <bool MERGED = false>
__device__ __forceinline__ void ldlt_diag_factor(double* S, double* X, double* misc, int b, int ld,
                                                 const int32_t* __restrict__ Nvec, int Nfixed, int k,
                                                 double* __restrict__ K, double* __restrict__ dvec,
                                                 int32_t* __restrict__ info, int32_t* __restrict__ nneg,
                                                 const int32_t* __restrict__ npos_expected) {
    constexpr int WP = MERGED ? WNPM : WNP;  // pitch of the -W panel
    const int j0 = k * NB;
    double* Wn = misc;                       // 64 x WP : -W of the current panel
    double* T = Wn + NB * WNP;               // 32 x TP (not MERGED)
    double* rinv = MERGED ? Wn + NB * WNPM : T + 32 * TP;  // 64
    int* flags = reinterpret_cast<int*>(rinv + NB);
    int& s_bad = flags[0];
    int& s_neg = flags[1];
    int& s_sign = flags[2];
    double* Kb = K + (size_t)b * ld * ld;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, g = lane >> 2, q = lane & 3;
    const int N = Nvec != nullptr ? Nvec[b] : Nfixed;
    const int npos = npos_expected != nullptr ? npos_expected[b] : -1;
    if (tid == 0) { s_bad = 0x7fffffff; s_neg = 0; s_sign = 0; }
    TRACE_BEGIN(k, 0);

    // ---- factorisation
    double p0 = 0.0, p1 = 0.0;  // warp 0: pivot block, lane (g, q) owns P[g][2q], P[g][2q + 1]
    if (wid == 0) {
        const double2 v = *reinterpret_cast<const double2*>(S + g * DP + 2 * q);
        p0 = v.x;
        p1 = v.y;
    }
    __syncthreads();  // X zeroed, flags initialised
    FINE_DECL
    FINE_MARK();
    // X = inv(L) (unit lower) by block rows, behind warp 0's pivot chain: once Y_i = inv(L_ii) and the block row L[i, 0..i) are
    // known, X[i][c] = -Y_i sum_{c <= k < i} L[i][k] X[k][c] for c < i.  Warps 1..7 (one 8 x 8 tile each) compute block row
    // jb - 1 while warp 0 eliminates pivot block jb -- they have nothing else to do there -- so that only the last block row
    // is left after the loop (the block recursion 8 -> 16 -> 32 -> 64 after the factorisation cost 7-16 k cycles per diagonal block).
    // No tile above the block diagonal of X is ever read, so X needs no initialisation.
    auto x_row_tile = [&](int i, int c) {
        double t0 = 0.0, t1 = 0.0;
        for (int k = c; k < i; k++) {
            const double* ap = S + (i * 8 + g) * DP + k * 8 + q;
            if constexpr (!MERGED) {
                const double* bp = X + (k * 8 + q) * DP + c * 8 + g;
                dmma884(t0, t1, ap[0], bp[0]);
                dmma884(t0, t1, ap[4], bp[4 * DP]);
            } else {
                // B[kk][n] = X[8k + kk][8c + n] sits at S[8c + n][8k + kk]; k == c: the unit lower 8 x 8 block Y_c
                const double* bp = S + (c * 8 + g) * DP + k * 8 + q;
                double v0 = bp[0], v1 = bp[4];
                if (k == c) {
                    v0 = q > g ? v0 : (q == g ? 1.0 : 0.0);
                    v1 = q + 4 > g ? v1 : (q + 4 == g ? 1.0 : 0.0);
                }
                dmma884(t0, t1, ap[0], v0);
                dmma884(t0, t1, ap[4], v1);
            }
        }
        double x0 = 0.0, x1 = 0.0;
        if constexpr (!MERGED) {
            double* scr = T + wid * 64;  // the 8 x 8 product as the B operand of the second product
            *reinterpret_cast<double2*>(scr + g * 8 + 2 * q) = make_double2(t0, t1);
            __syncwarp();
            const double* ya = X + (i * 8 + g) * DP + i * 8 + q;
            dmma884(x0, x1, ya[0], scr[q * 8 + g]);
            dmma884(x0, x1, ya[4], scr[(q + 4) * 8 + g]);
            __syncwarp();
            *reinterpret_cast<double2*>(X + (i * 8 + g) * DP + c * 8 + 2 * q) = make_double2(dneg(x0), dneg(x1));
        } else {
            // T[r][n] is held by lane (r, n / 2), slot n & 1; this lane's B fragment is T[q][g], T[q + 4][g]
            const int s0 = q * 4 + (g >> 1), s1 = (q + 4) * 4 + (g >> 1);
            const double ta0 = __shfl_sync(0xffffffffu, t0, s0), ta1 = __shfl_sync(0xffffffffu, t1, s0);
            const double tb0 = __shfl_sync(0xffffffffu, t0, s1), tb1 = __shfl_sync(0xffffffffu, t1, s1);
            const double b0 = (g & 1) ? ta1 : ta0, b1 = (g & 1) ? tb1 : tb0;
            // A[g][kk] = Y_i[g][kk] sits at S[8i + kk][8i + g] for kk < g
            const double* yp = S + (i * 8 + q) * DP + i * 8 + g;
            const double a0 = q < g ? yp[0] : (q == g ? 1.0 : 0.0);
            const double a1 = q + 4 < g ? yp[4 * DP] : (q + 4 == g ? 1.0 : 0.0);
            dmma884(x0, x1, a0, b0);
            dmma884(x0, x1, a1, b1);
            S[(c * 8 + 2 * q) * DP + i * 8 + g] = dneg(x0);      // X[8i + g][8c + 2q] transposed
            S[(c * 8 + 2 * q + 1) * DP + i * 8 + g] = dneg(x1);
        }
    };
    for (int jb = 0; jb < 8; jb++) {
        const int c0 = jb * 8;
        if (wid != 0 && wid < jb) x_row_tile(jb - 1, wid - 1);  // block row jb - 1: tiles c = 0 .. jb - 2
        if (wid == 0) {
            // eliminate the 8 x 8 pivot block in registers
            // Every rank-1 update of a pivot step is ONE DMMA whose operands sit in the registers of the lanes that
            // already own them: column j of P lives in the lanes with q == j / 2 (slot j & 1), which are exactly the lanes
            // that supply k-slot j / 2 of the A fragment (rows) and of the B fragment (columns) -- no exchange at all.
            // What is left per pivot is one broadcast of d_j, the reciprocal and the scaling of the column; the shuffle
            // version of round 1 (six 64-bit shuffles per pivot, scalar updates) ran at 275 cycles per pivot (tools/ldlt_trace).
            // Y = inv(L_pivot) is updated the same way; its B operand (row j of Y) comes from Z = Y', kept alongside.
            double y0 = (2 * q == g) ? 1.0 : 0.0, y1 = (2 * q + 1 == g) ? 1.0 : 0.0;
            double z0 = y0, z1 = y1;
            double rc0 = 0.0, rc1 = 0.0;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const bool mine = q == (j >> 1);
                const double pj = (j & 1) ? p1 : p0;                 // P[g][j] in the lanes with q == j / 2
                const double dj = __shfl_sync(0xffffffffu, pj, j * 4 + (j >> 1));
                double r0;
                asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(dj));
                if (dj == 0.0) r0 = 0.0;
                const double w = (mine && g > j) ? pj : 0.0;         // column j below the pivot, zero elsewhere
                const double t = w * r0;                             // beside e -> sc on the way to 1 / dj
                const double e = fma(-dj, r0, 1.0);
                const double sc = fma(e, e, e);                      // 1 / dj = r0 (1 + e + e^2): cubic in the seed's error
                const double l = fma(t, sc, t);                      // L[g][j]
                const double rj = fma(r0, sc, r0);
                if (j == 2 * q) rc0 = rj;
                if (j == 2 * q + 1) rc1 = rj;
                dmma884(p0, p1, dneg(w), l);                         // P[g][n] -= w_g l_n   (g, n > j)
                const double zj = mine ? ((j & 1) ? z1 : z0) : 0.0;  // Z[n][j] = Y[j][n]
                dmma884(y0, y1, dneg(l), zj);                        // Y[g][n] -= l_g Y[j][n]
                dmma884(z0, z1, dneg(zj), l);                        // Z[n][g] -= Y[j][n] l_g
                // column j is final: keep L[g][j] itself (the value the update used) in its slot
                if (mine && g > j) { if (j & 1) p1 = l; else p0 = l; }
            }
            // pivots: diagnostics, d, 1/d
            if ((g >> 1) == q) {
                const double d = (g & 1) ? p1 : p0;
                const int jl = c0 + g, j = j0 + jl;
                if (!(isfinite(d)) || d == 0.0) atomicMin(&s_bad, jl + 1);
                if (d < 0.0 && j < N) atomicAdd(&s_neg, 1);
                // quasi-definite sign pattern: the first npos pivots positive, the remaining (up to N) negative
                if (npos >= 0 && j < N && ((j < npos) != (d > 0.0))) s_sign = 1;
                dvec[(size_t)b * ld + j] = d;
                rinv[jl] = (g & 1) ? rc1 : rc0;
            }
            // L (strict lower, scaled), d on the diagonal; Y = inv(L_pivot) into the diagonal block of X
            const double l0 = p0, l1 = p1;
            if (2 * q + 1 <= g) *reinterpret_cast<double2*>(S + (c0 + g) * DP + c0 + 2 * q) = make_double2(l0, l1);
            else if (2 * q == g) S[(c0 + g) * DP + c0 + 2 * q] = l0;
            if constexpr (!MERGED) {
                *reinterpret_cast<double2*>(X + (c0 + g) * DP + c0 + 2 * q) = make_double2(y0, y1);
            } else {  // strictly lower part of Y, transposed into the upper triangle of the pivot block
                if (2 * q < g) S[(c0 + 2 * q) * DP + c0 + g] = y0;
                if (2 * q + 1 < g) S[(c0 + 2 * q + 1) * DP + c0 + g] = y1;
            }
        }
        FINE_MARK();  // [0] pivot block
        __syncthreads();  // pivot block published; all trailing updates of the previous panel are done
        FINE_MARK();  // [1] barrier
        if (jb == 7) break;
        // panel below the pivot block: row tile t = jb + 1 + wid
        {
            const int t = jb + 1 + wid;
            if (t < 8) {
                const double* ap = S + (t * 8 + g) * DP + c0 + q;
                double w0 = 0.0, w1 = 0.0;
                const double a0 = ap[0], a1 = ap[4];
                double b0, b1;  // B[kk][n] = Y[n][kk], kk = q, q + 4, n = g
                if constexpr (!MERGED) {
                    const double* bp = X + (c0 + g) * DP + c0 + q;
                    b0 = bp[0];
                    b1 = bp[4];
                } else {
                    const double* bp = S + (c0 + q) * DP + c0 + g;  // Y[n][kk] at S[c0 + kk][c0 + n] for n > kk
                    b0 = g > q ? bp[0] : (g == q ? 1.0 : 0.0);
                    b1 = g > q + 4 ? bp[4 * DP] : (g == q + 4 ? 1.0 : 0.0);
                }
                dmma884(w0, w1, a0, b0);
                dmma884(w0, w1, a1, b1);
                __syncwarp();
                const double2 rv = *reinterpret_cast<const double2*>(rinv + c0 + 2 * q);
                *reinterpret_cast<double2*>(S + (t * 8 + g) * DP + c0 + 2 * q) = make_double2(w0 * rv.x, w1 * rv.y);
                *reinterpret_cast<double2*>(Wn + (t * 8 + g) * WP + 2 * q) = make_double2(dneg(w0), dneg(w1));
            }
        }
        FINE_MARK();  // [2] panel
        __syncthreads();
        FINE_MARK();  // [3] barrier
        // trailing update by 8 x 8 tiles (ti, tj), jb < tj <= ti < 8: warp 0 takes the next pivot block and keeps
        // it in registers, warps 1..7 share the rest
        {
            const int m = 7 - jb;                      // tile rows / cols left
            const int ntile = m * (m + 1) / 2;
            for (int e = (wid == 0 ? 0 : wid); e < ntile; e += (wid == 0 ? ntile : 7)) {
                // e -> (ti, tj) in row-major order of the lower triangle: e = ti' (ti' + 1) / 2 + tj'
                int tr = 0;
                while ((tr + 1) * (tr + 2) / 2 <= e) tr++;
                const int tc = e - tr * (tr + 1) / 2;
                const int ti = jb + 1 + tr, tj = jb + 1 + tc;
                double* cp = S + (ti * 8 + g) * DP + tj * 8 + 2 * q;
                double2 cv = *reinterpret_cast<const double2*>(cp);
                const double* ap = Wn + (ti * 8 + g) * WP + q;
                const double* bp = S + (tj * 8 + g) * DP + c0 + q;
                dmma884(cv.x, cv.y, ap[0], bp[0]);
                dmma884(cv.x, cv.y, ap[4], bp[4]);
                if (wid == 0) { p0 = cv.x; p1 = cv.y; }
                else *reinterpret_cast<double2*>(cp) = cv;
            }
        }
        // (no barrier: warp 0 goes on with the pivot block in registers; the others wait at the next barrier)
        FINE_MARK();  // [4] update (warp 0: the next pivot block only)
    }
    FINE_FLUSH();
    TRACE_MARK(8);

    // ---- last block row of X = inv(L) (the loop left at its barrier after pivot block 7)
    if (wid < 7) x_row_tile(7, wid);
    __syncthreads();
    TRACE_MARK(9);

    // ---- write back: L (strict lower), d (diagonal), inv(L)' (strict upper)
    for (int e = tid; e < NB * NB; e += blockDim.x) {
        const int r = e >> 6, c = e & 63;
        const double v = (MERGED || c <= r) ? S[r * DP + c] : X[c * DP + r];  // K[j0 + r][j0 + c] = X[c][r], c > r
        Kb[(size_t)(j0 + r) * ld + j0 + c] = v;
    }
    if (tid == 0) {
        int bad = 0;
        if (s_bad != 0x7fffffff) bad = j0 + s_bad;        // zero / non-finite pivot at column `bad`
        else if (s_sign) bad = GF_INFO_NOT_QUASIDEFINITE;  // wrong pivot sign: unpivoted LDL' not trusted
        if (k == 0) { nneg[b] = s_neg; info[b] = bad; }
        else {
            nneg[b] += s_neg;
            if (bad != 0 && info[b] == 0) info[b] = bad;
        }
    }
    TRACE_MARK(10);
    TRACE_COUNT();
}

// ------------------------------------------------------------------------------------------------
constexpr int TM = 128, TN = 64, STAGES = 2;
constexpr int PKC = 32;        // k-chunk per pipeline stage
constexpr int PSP = PKC + 4;   // its smem row pitch (== 4 mod 16: conflict-free fragment loads)
constexpr int STAGE_ROWS = TM + TN;  // rows of a stage: panel 128 A + 64 B, chain 64 A + 64 B1 + 64 B2
constexpr int PN_SMEM = STAGES * STAGE_ROWS * PSP * (int)sizeof(double);
static_assert(PKC * 2 == NB, "two chunks per block column: the chunk count is even, the last chunk sits in stage 1");
// epilogue regions (doubles from the start of the dynamic shared memory); Xs must fit into stage 0, which is idle
// while the last chunk (always in stage 1) is computed and receives the diagonal block of column k meanwhile
constexpr int EPI_XS = 0;                       // 64 x EP : storage rows of diagonal block k (inv(L_kk)' above its diagonal)
static_assert(NB * EP <= STAGE_ROWS * PSP, "Xs must fit into one pipeline stage");
constexpr int PN_CS = NB * EP;                  // panel: C tile, 128 x EP (written after the main loop)
constexpr int PN_RINV = PN_CS + TM * EP;        // 64 reciprocal pivots
static_assert((PN_RINV + NB) * (int)sizeof(double) <= PN_SMEM, "panel epilogue must fit");
constexpr int CH_S = NB * EP;                   // chain: S (64 x DP), C / W' (64 x EP), rinv; factor scratch over C
constexpr int CH_CS = CH_S + DG_S;
constexpr int CH_RINV = CH_CS + NB * EP;
static_assert((CH_RINV + NB) * (int)sizeof(double) <= PN_SMEM, "chain epilogue must fit");
static_assert(DG_MISC <= NB * EP, "the factor's scratch reuses the C / W' region");
static_assert(DG_OPS <= NB * EP, "inv(L) scratch reuses the Xs region");

// 16-byte asynchronous copy of the 64 x 64 diagonal block of column k (storage layout) into Xs
__device__ __forceinline__ void load_diag_block(double* Xs, const double* __restrict__ Kb, int j0, int ld) {
#pragma unroll
    for (int t = 0; t < 8; t++) {
        const int piece = threadIdx.x + t * 256;
        const int r = piece >> 5, c2 = (piece & 31) * 2;
        cp_async16(Xs + r * EP + c2, Kb + (size_t)(j0 + r) * ld + j0 + c2);
    }
}

// B fragment of X' for the triangular solve, read from the storage layout G[kk][n] (X[n][kk] for n > kk):
// unit diagonal and zeros below it are supplied here.
__device__ __forceinline__ double xt_fragment(const double* Xs, int kq, int n) {
    const double v = Xs[kq * EP + n];
    return n > kq ? v : (n == kq ? 1.0 : 0.0);
}

// One ROWS x 64 tile of block column k (rows i0.., all inside the padded order): update + triangular solve.
// ROWS = 128: 4 x 2 warps of 32 x 32;  ROWS = 64 (odd remainder block): 2 x 4 warps of 32 x 16.
// LROWS: A rows the shared-memory layout is sized for (128 inside the column kernel, 64 in the 64-row-only kernel).
template <int ROWS, int LROWS = TM>
__device__ __forceinline__ void ldlt_panel_tile(double* sm, int i0, int j0, int ld, double* __restrict__ Kb,
                                                const double* db, const KktView& kv) {
    constexpr int WN = (ROWS == 128) ? 2 : 4;   // warps along the 64 columns
    constexpr int NI = 8 / WN;                  // 8-column DMMA tiles per warp
    constexpr int WC = NI * 8;                  // columns per warp
    constexpr int AT = ROWS / 16;               // 16-byte pieces per thread per stage for the A rows
    double* As = sm;                            // stage s: A rows at s * STAGE_ROWS * PSP, B rows TM rows later
    constexpr int SSTRIDE = (LROWS + TN) * PSP;  // doubles per pipeline stage
    static_assert(NB * EP <= SSTRIDE, "Xs must fit into one pipeline stage");
    static_assert(ROWS <= LROWS, "layout too small");
    double* Bs = sm + LROWS * PSP;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int wm = wid / WN, wn = wid % WN;
    const int g = lane >> 2, q = lane & 3;
    const int nchunks = j0 / PKC;
    TRACE_BEGIN(j0 / NB, 1);

    // Each thread copies one 16-byte piece of AT A rows and 4 B rows per stage: row = (tid >> 4) + 16 t,
    // piece = tid & 15.  A = W'[tile rows, chunk] from the mirrored blocks (p, i): tile row r of block i, depth
    // 64 p + c sits at K[64 p + (r & 63)][64 i + c];  B = L[block k rows, chunk] from the lower triangle.
    const int lrow = tid >> 4, lpart = (tid & 15) * 2;
    const double* gA = Kb + (size_t)lrow * ld + i0 + lpart;
    const double* gB = Kb + (size_t)(j0 + lrow) * ld + lpart;
    const size_t gstride = (size_t)16 * ld;
    double* sA = As + lrow * PSP + lpart;
    double* sB = Bs + lrow * PSP + lpart;
    auto load_stage = [&](int chunk, int stage) {
        const double* ga = gA + (size_t)(chunk >> 1) * NB * ld + (chunk & 1) * PKC;
        const double* gb = gB + chunk * PKC;
        double* sa = sA + stage * SSTRIDE;
        double* sb = sB + stage * SSTRIDE;
#pragma unroll
        for (int t = 0; t < AT; t++) cp_async16(sa + t * 16 * PSP, ga + (t & 3) * gstride + (t >> 2) * NB);
#pragma unroll
        for (int t = 0; t < 4; t++) cp_async16(sb + t * 16 * PSP, gb + t * gstride);
    };
    if (nchunks > 0) load_stage(0, 0);
    else load_diag_block(sm + EPI_XS, Kb, j0, ld);
    cp_async_commit();
    // reciprocal pivots of block column k for the epilogue: fetched and inverted now, behind the main loop
    const double my_rinv = (tid < NB) ? 1.0 / db[j0 + tid] : 0.0;  // plain load: the whole-matrix kernel wrote it
    // accumulators start from the current A tile (these loads overlap the first pipeline stage)
    double acc[4][NI][2];
#pragma unroll
    for (int mi = 0; mi < 4; mi++) {
        const int rr = i0 + wm * 32 + mi * 8 + g, cc = j0 + wn * WC + 2 * q;
#pragma unroll
        for (int ni = 0; ni < NI; ni++) {
            const double2 v = kkt_load2(kv, Kb, ld, rr, cc + ni * 8);
            acc[mi][ni][0] = v.x;
            acc[mi][ni][1] = v.y;
        }
    }
    TRACE_MARK(0);
    for (int c = 0; c < nchunks; c++) {
        cp_async_wait<0>();
        __syncthreads();
        if (c + 1 < nchunks) load_stage(c + 1, (c + 1) & 1);
        else load_diag_block(sm + EPI_XS, Kb, j0, ld);   // stage 0 is idle during the last chunk
        cp_async_commit();
        const double* as = As + (c & 1) * SSTRIDE + (wm * 32 + g) * PSP + q;
        const double* bs = Bs + (c & 1) * SSTRIDE + (wn * WC + g) * PSP + q;
#pragma unroll
        for (int kk = 0; kk < PKC; kk += 4) {
            double a[4], bf[NI];
#pragma unroll
            for (int mi = 0; mi < 4; mi++) a[mi] = as[mi * 8 * PSP + kk];
#pragma unroll
            for (int ni = 0; ni < NI; ni++) bf[ni] = bs[ni * 8 * PSP + kk];
#pragma unroll
            for (int mi = 0; mi < 4; mi++)
#pragma unroll
                for (int ni = 0; ni < NI; ni++) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], bf[ni]);
        }
    }
    cp_async_wait<0>();
    __syncthreads();  // every warp is done with the pipeline buffers; the diagonal block has landed in Xs
    TRACE_MARK(1);

    // ---- epilogue: W = C X' (DMMA), L[i,k] = W D^{-1} -> lower triangle, W' = -W -> mirrored block
    const double* Xs = sm + EPI_XS;
    double* Cs = sm + PN_CS;
    double* rinv = sm + PN_CS + LROWS * EP;
    static_assert((PN_CS + LROWS * EP + NB) <= STAGES * SSTRIDE, "panel epilogue must fit");
#pragma unroll
    for (int mi = 0; mi < 4; mi++) {
        const int r = wm * 32 + mi * 8 + g;
#pragma unroll
        for (int ni = 0; ni < NI; ni++) {
            const int c = wn * WC + ni * 8 + 2 * q;
            *reinterpret_cast<double2*>(Cs + r * EP + c) = make_double2(acc[mi][ni][0], acc[mi][ni][1]);
        }
    }
    if (tid < NB) rinv[tid] = my_rinv;
    __syncthreads();
    TRACE_MARK(2);
    // every warp solves ROWS / 8 full rows (all 64 columns), so the triangular pruning of the k range loads all
    // warps -- and with them the four SM sub-partitions -- equally
    {
        constexpr int TMI = ROWS / 64;           // 8-row DMMA tiles per warp
        constexpr int WR = ROWS / 8;             // rows per warp
        double t[TMI][8][2];
#pragma unroll
        for (int mi = 0; mi < TMI; mi++)
#pragma unroll
            for (int ni = 0; ni < 8; ni++) { t[mi][ni][0] = 0.0; t[mi][ni][1] = 0.0; }
        const double* as = Cs + (wid * WR + g) * EP + q;
#pragma unroll
        for (int kk = 0; kk < NB; kk += 4) {
            double a[TMI];
#pragma unroll
            for (int mi = 0; mi < TMI; mi++) a[mi] = as[mi * 8 * EP + kk];
#pragma unroll
            for (int ni = 0; ni < 8; ni++) {
                if (kk < ni * 8 + 8) {  // X[n][kk] = 0 for kk > n
                    const double bf = xt_fragment(Xs, kk + q, ni * 8 + g);
#pragma unroll
                    for (int mi = 0; mi < TMI; mi++) dmma884(t[mi][ni][0], t[mi][ni][1], a[mi], bf);
                }
            }
        }
#pragma unroll
        for (int ni = 0; ni < 8; ni++) {
            const int c = ni * 8 + 2 * q;
            const double r0 = rinv[c], r1 = rinv[c + 1];
#pragma unroll
            for (int mi = 0; mi < TMI; mi++) {
                const int r = wid * WR + mi * 8 + g;
                const double w0 = t[mi][ni][0], w1 = t[mi][ni][1];
                *reinterpret_cast<double2*>(Kb + (size_t)(i0 + r) * ld + j0 + c) = make_double2(w0 * r0, w1 * r1);
                *reinterpret_cast<double2*>(Kb + (size_t)(j0 + (r & 63)) * ld + i0 + (r >> 6) * NB + c) =
                    make_double2(-w0, -w1);
            }
        }
    }
    TRACE_MARK(3);
    TRACE_COUNT();
}

// ------------------------------------------------------------------------------------------------
// Chain CTA of block column k: rows of block k+1 (i0 = j0 + 64).  One DMMA main loop over the block columns
// p < k produces the 64 x 128 tile [ A[k+1,k] | A[k+1,k+1] ] + sum_p W'[k+1,p] [ L[k,p] | L[k+1,p] ]'
// (8 warps of 32 x 32), then
//   W = C X_k' (DMMA), L[k+1,k] = W D_k^{-1} -> global, W' = -W -> mirror;  S = A[k+1,k+1]-part + W' L[k+1,k]';
//   ldlt_diag_factor(S) for block k+1.
__device__ __forceinline__ void ldlt_chain_body(double* sm, int b, int ld, const int32_t* __restrict__ Nvec,
                                                int Nfixed, int k, double* __restrict__ K,
                                                double* __restrict__ dvec, int32_t* __restrict__ info,
                                                int32_t* __restrict__ nneg,
                                                const int32_t* __restrict__ npos_expected, const KktView& kv) {
    const int Np = padded_order(Nvec, Nfixed, b, ld);
    const int j0 = k * NB, i0 = j0 + NB;
    if (i0 >= Np) return;
    double* Kb = K + (size_t)b * ld * ld;
    const double* db = dvec + (size_t)b * ld;
    double* As = sm;                    // per stage: 64 rows W'[k+1, chunk], 64 rows L[k, chunk], 64 rows L[k+1, chunk]
    double* B1s = sm + NB * PSP;
    double* B2s = sm + 2 * NB * PSP;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int wm = wid >> 2, wn = wid & 3;      // 2 x 4 warps of 32 x 32
    const int g = lane >> 2, q = lane & 3;
    const bool diag_part = wn >= 2;             // columns 64..127 of the tile = diagonal tile of block k+1
    const bool skip = (wm == 0) && (wn == 3);   // its strictly upper 32 x 32 quadrant is never read
    const int nchunks = j0 / PKC;
    TRACE_BEGIN(k + 1, 0);

    const int lrow = tid >> 4, lpart = (tid & 15) * 2;
    const double* gA = Kb + (size_t)lrow * ld + i0 + lpart;
    const double* gB1 = Kb + (size_t)(j0 + lrow) * ld + lpart;
    const double* gB2 = Kb + (size_t)(i0 + lrow) * ld + lpart;
    const size_t gstride = (size_t)16 * ld;
    auto load_stage = [&](int chunk, int stage) {
        const double* ga = gA + (size_t)(chunk >> 1) * NB * ld + (chunk & 1) * PKC;
        const double* gb1 = gB1 + chunk * PKC;
        const double* gb2 = gB2 + chunk * PKC;
        double* sa = As + stage * STAGE_ROWS * PSP + lrow * PSP + lpart;
#pragma unroll
        for (int t = 0; t < 4; t++) cp_async16(sa + t * 16 * PSP, ga + t * gstride);
#pragma unroll
        for (int t = 0; t < 4; t++) cp_async16(sa + (NB + t * 16) * PSP, gb1 + t * gstride);
#pragma unroll
        for (int t = 0; t < 4; t++) cp_async16(sa + (2 * NB + t * 16) * PSP, gb2 + t * gstride);
    };
    if (nchunks > 0) load_stage(0, 0);
    else load_diag_block(sm + EPI_XS, Kb, j0, ld);
    cp_async_commit();
    const double my_rinv = (tid < NB) ? 1.0 / db[j0 + tid] : 0.0;  // for the epilogue, behind the main loop
    double acc[4][4][2];
    {
        const int colbase = (diag_part ? i0 + (wn - 2) * 32 : j0 + wn * 32) + 2 * q;
#pragma unroll
        for (int mi = 0; mi < 4; mi++) {
            const int rr = i0 + wm * 32 + mi * 8 + g;
#pragma unroll
            for (int ni = 0; ni < 4; ni++) {
                double2 v = make_double2(0.0, 0.0);
                if (!skip) v = kkt_load2(kv, Kb, ld, rr, colbase + ni * 8);
                acc[mi][ni][0] = v.x;
                acc[mi][ni][1] = v.y;
            }
        }
    }
    TRACE_MARK(0);
    for (int c = 0; c < nchunks; c++) {
        cp_async_wait<0>();
        __syncthreads();
        if (c + 1 < nchunks) load_stage(c + 1, (c + 1) & 1);
        else load_diag_block(sm + EPI_XS, Kb, j0, ld);
        cp_async_commit();
        const int so = (c & 1) * STAGE_ROWS * PSP;
        const double* as = As + so + (wm * 32 + g) * PSP + q;
        const double* bs = (diag_part ? B2s + so + ((wn - 2) * 32 + g) * PSP : B1s + so + (wn * 32 + g) * PSP) + q;
        if (!skip) {
#pragma unroll
            for (int kk = 0; kk < PKC; kk += 4) {
                double a[4], bf[4];
#pragma unroll
                for (int mi = 0; mi < 4; mi++) a[mi] = as[mi * 8 * PSP + kk];
#pragma unroll
                for (int ni = 0; ni < 4; ni++) bf[ni] = bs[ni * 8 * PSP + kk];
#pragma unroll
                for (int mi = 0; mi < 4; mi++)
#pragma unroll
                    for (int ni = 0; ni < 4; ni++) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], bf[ni]);
            }
        }
    }
    cp_async_wait<0>();
    __syncthreads();  // every warp is done with the pipeline buffers; the diagonal block has landed in Xs
    TRACE_MARK(1);

    // ---- epilogue 1: panel half -> Cs, diagonal half -> S
    double* Xs = sm + EPI_XS;                                          // 64 x EP, later L[k+1,k]
    double(*S)[DP] = reinterpret_cast<double(*)[DP]>(sm + CH_S);           // 64 x DP
    double* Cs = sm + CH_CS;                                           // 64 x EP, later W' = -C X'
    double* rinv = sm + CH_RINV;                                       // 64
    if (!diag_part) {
#pragma unroll
        for (int mi = 0; mi < 4; mi++) {
            const int r = wm * 32 + mi * 8 + g;
#pragma unroll
            for (int ni = 0; ni < 4; ni++) {
                const int c = wn * 32 + ni * 8 + 2 * q;
                *reinterpret_cast<double2*>(Cs + r * EP + c) = make_double2(acc[mi][ni][0], acc[mi][ni][1]);
            }
        }
    } else if (!skip) {
#pragma unroll
        for (int mi = 0; mi < 4; mi++) {
            const int r = wm * 32 + mi * 8 + g;
#pragma unroll
            for (int ni = 0; ni < 4; ni++) {
                const int c = (wn - 2) * 32 + ni * 8 + 2 * q;
                S[r][c] = acc[mi][ni][0];
                S[r][c + 1] = acc[mi][ni][1];
            }
        }
    }
    if (tid < NB) rinv[tid] = my_rinv;
    __syncthreads();
    TRACE_MARK(2);

    // ---- epilogue 2: W = C X' (every warp 8 full rows), L = W D^{-1}
    {
        double t[8][2];
#pragma unroll
        for (int ni = 0; ni < 8; ni++) { t[ni][0] = 0.0; t[ni][1] = 0.0; }
        const double* as = Cs + (wid * 8 + g) * EP + q;
#pragma unroll
        for (int kk = 0; kk < NB; kk += 4) {
            const double a = as[kk];
#pragma unroll
            for (int ni = 0; ni < 8; ni++) {
                if (kk < ni * 8 + 8) dmma884(t[ni][0], t[ni][1], a, xt_fragment(Xs, kk + q, ni * 8 + g));
            }
        }
        __syncthreads();  // all reads of Cs / Xs are done: overwrite them with W' and L
        const int r = wid * 8 + g;
#pragma unroll
        for (int ni = 0; ni < 8; ni++) {
            const int c = ni * 8 + 2 * q;
            const double r0 = rinv[c], r1 = rinv[c + 1];
            const double2 wv = make_double2(-t[ni][0], -t[ni][1]);
            const double2 lv = make_double2(t[ni][0] * r0, t[ni][1] * r1);
            *reinterpret_cast<double2*>(Cs + r * EP + c) = wv;
            *reinterpret_cast<double2*>(Xs + r * EP + c) = lv;
            *reinterpret_cast<double2*>(Kb + (size_t)(i0 + r) * ld + j0 + c) = lv;
            *reinterpret_cast<double2*>(Kb + (size_t)(j0 + r) * ld + i0 + c) = wv;
        }
    }
    __syncthreads();
    TRACE_MARK(3);

    // ---- epilogue 3: S += W' L'  (4 x 2 warps of 16 x 32, strictly upper quadrant skipped)
    {
        const int wm2 = wid >> 1, wn2 = wid & 1;
        if (!((wn2 == 1) && (wm2 < 2))) {
            double s2[2][4][2];
#pragma unroll
            for (int mi = 0; mi < 2; mi++) {
                const int r = wm2 * 16 + mi * 8 + g;
#pragma unroll
                for (int ni = 0; ni < 4; ni++) {
                    const int c = wn2 * 32 + ni * 8 + 2 * q;
                    s2[mi][ni][0] = S[r][c];
                    s2[mi][ni][1] = S[r][c + 1];
                }
            }
            const double* as = Cs + (wm2 * 16 + g) * EP + q;
            const double* bs = Xs + (wn2 * 32 + g) * EP + q;
#pragma unroll 4
            for (int kk = 0; kk < NB; kk += 4) {
                double a[2], bf[4];
#pragma unroll
                for (int mi = 0; mi < 2; mi++) a[mi] = as[mi * 8 * EP + kk];
#pragma unroll
                for (int ni = 0; ni < 4; ni++) bf[ni] = bs[ni * 8 * EP + kk];
#pragma unroll
                for (int mi = 0; mi < 2; mi++)
#pragma unroll
                    for (int ni = 0; ni < 4; ni++) dmma884(s2[mi][ni][0], s2[mi][ni][1], a[mi], bf[ni]);
            }
#pragma unroll
            for (int mi = 0; mi < 2; mi++) {
                const int r = wm2 * 16 + mi * 8 + g;
#pragma unroll
                for (int ni = 0; ni < 4; ni++) {
                    const int c = wn2 * 32 + ni * 8 + 2 * q;
                    S[r][c] = s2[mi][ni][0];
                    S[r][c + 1] = s2[mi][ni][1];
                }
            }
        }
    }
    __syncthreads();
    TRACE_MARK(4);
    ldlt_diag_factor(sm + CH_S, sm + EPI_XS, sm + CH_CS, b, ld, Nvec, Nfixed, k + 1, K, dvec, info, nneg,
                     npos_expected);
}

// ------------------------------------------------------------------------------------------------
// Launch wrappers.
__global__ void __launch_bounds__(256, 3) ldlt_diag0_kernel(int ld, const int32_t* __restrict__ Nvec, int Nfixed,
                                                             double* __restrict__ K, double* __restrict__ dvec,
                                                             int32_t* __restrict__ info, int32_t* __restrict__ nneg,
                                                             const int32_t* __restrict__ npos_expected, GfWork work,
                                                             int woff, KktSrc src) {
    const int b = gf_instance(work, woff + blockIdx.x);
    if (b < 0) return;
    extern __shared__ double sm[];
    const KktView kv = kkt_view(src, b);
    if (padded_order(Nvec, Nfixed, b, ld) <= 0) {  // empty system (everything active, no constraints): trivially ok
        if (threadIdx.x == 0) { info[b] = 0; nneg[b] = 0; }
        return;
    }
    double(*S)[DP] = reinterpret_cast<double(*)[DP]>(sm);
    const double* Kb = K + (size_t)b * ld * ld;
    for (int e = threadIdx.x; e < NB * NB / 2; e += blockDim.x) {
        const int r = e >> 5, c = (e & 31) * 2;
        const double2 v = kkt_load2(kv, Kb, ld, r, c);
        S[r][c] = v.x;
        S[r][c + 1] = v.y;
    }
    __syncthreads();
    ldlt_diag_factor(sm, sm + DG_S, sm + DG_S + DG_OPS, b, ld, Nvec, Nfixed, 0, K, dvec, info, nneg, npos_expected);
}

// Block column k: blockIdx.x == 0 is the chain CTA (rows of block k+1 + diagonal block k+1), blockIdx.x >= 1 the
// 128-row panel tiles from row j0 + 128 on.  No CTA depends on another CTA of the same launch.
__global__ void __launch_bounds__(256, 2) ldlt_column_kernel(int ld, const int32_t* __restrict__ Nvec, int Nfixed,
                                                             int k, double* __restrict__ K, double* __restrict__ dvec,
                                                             int32_t* __restrict__ info, int32_t* __restrict__ nneg,
                                                             const int32_t* __restrict__ npos_expected, GfWork work,
                                                             int woff, int cnt, int tiles, KktSrc src, int order) {
    // 1-D grid of cnt * tiles CTAs: the cnt chain CTAs (the long ones) come first, then the panel tiles
    // (order 1 / 2, developer experiments GF_LDLT_ORDER: matrix by matrix / panel tiles first)
    const int lin = blockIdx.x;
    int wi, tile;
    if (order == 0) {
        wi = lin < cnt ? lin : (lin - cnt) / (tiles - 1);
        tile = lin < cnt ? 0 : 1 + (lin - cnt) % (tiles - 1);
    } else if (order == 1) {
        wi = lin / tiles;
        tile = lin % tiles;
    } else {
        const int np = cnt * (tiles - 1);
        wi = lin < np ? lin / (tiles - 1) : lin - np;
        tile = lin < np ? 1 + lin % (tiles - 1) : 0;
    }
    const int b = gf_instance(work, woff + wi);
    if (b < 0) return;
    extern __shared__ double sm[];
    const KktView kv = kkt_view(src, b);
    if (tile == 0) {
        ldlt_chain_body(sm, b, ld, Nvec, Nfixed, k, K, dvec, info, nneg, npos_expected, kv);
        return;
    }
    const int Np = padded_order(Nvec, Nfixed, b, ld);
    const int j0 = k * NB;
    const int i0 = j0 + 2 * NB + (tile - 1) * TM;
    if (i0 >= Np) return;
    double* Kb = K + (size_t)b * ld * ld;
    const double* db = dvec + (size_t)b * ld;
    if (Np - i0 >= TM) ldlt_panel_tile<128>(sm, i0, j0, ld, Kb, db, kv);
    else ldlt_panel_tile<64>(sm, i0, j0, ld, Kb, db, kv);   // odd remainder block: Np - i0 == 64
}

// Whole factorisation of one matrix by ONE CTA (opt-in experiment, GF_LDLT_WHOLE=1): diagonal block 0, then per block
// column k the chain tile (rows of block k+1 + the factorisation of diagonal block k+1) and the panel tiles below it,
// one after the other, with the device functions of the per-column kernels.  No launch boundaries, no per-column tails,
// and the two CTAs of an SM drift apart, so the latency-bound pivot chain of one runs beside the DMMA main loops of
// the other instead of beside another pivot chain.  Measured (B200, B = 4096): N = 768 27.5 ms vs 26.1 ms for the
// per-column launches, N = 512 10.0 vs 9.5, N = 1024 58.6 vs 55.7, N = 2048 (B = 1024) 102 vs 97 -- identical results,
// 5 % SLOWER: the pivot chain next to a DMMA stream is starved (tools/fp64_contention_bench.cu) for longer than the
// lock-step chain phases of the per-column launches leave the pipe idle.  Not adopted; kept for the record.
__global__ void __launch_bounds__(256, 2) ldlt_whole_kernel(int ld, const int32_t* __restrict__ Nvec, int Nfixed,
                                                            double* __restrict__ K, double* __restrict__ dvec,
                                                            int32_t* __restrict__ info, int32_t* __restrict__ nneg,
                                                            const int32_t* __restrict__ npos_expected, GfWork work,
                                                            int woff, KktSrc src) {
    const int b = gf_instance(work, woff + blockIdx.x);
    if (b < 0) return;
    extern __shared__ double sm[];
    const KktView kv = kkt_view(src, b);
    const int Np = padded_order(Nvec, Nfixed, b, ld);
    if (Np <= 0) {
        if (threadIdx.x == 0) { info[b] = 0; nneg[b] = 0; }
        return;
    }
    double* Kb = K + (size_t)b * ld * ld;
    {
        double(*S)[DP] = reinterpret_cast<double(*)[DP]>(sm);
        for (int e = threadIdx.x; e < NB * NB / 2; e += blockDim.x) {
            const int r = e >> 5, c = (e & 31) * 2;
            const double2 v = kkt_load2(kv, Kb, ld, r, c);
            S[r][c] = v.x;
            S[r][c + 1] = v.y;
        }
        __syncthreads();
        ldlt_diag_factor(sm, sm + DG_S, sm + DG_S + DG_OPS, b, ld, Nvec, Nfixed, 0, K, dvec, info, nneg, npos_expected);
    }
    const double* db = dvec + (size_t)b * ld;
    const int nblk = Np / NB;
    for (int k = 0; k + 1 < nblk; k++) {
        const int j0 = k * NB;
        __syncthreads();  // the previous phase's global writes (L, W', d) and shared-memory use are complete
        ldlt_chain_body(sm, b, ld, Nvec, Nfixed, k, K, dvec, info, nneg, npos_expected, kv);
        for (int i0 = j0 + 2 * NB; i0 < Np; i0 += TM) {
            __syncthreads();
            if (Np - i0 >= TM) ldlt_panel_tile<128>(sm, i0, j0, ld, Kb, db, kv);
            else ldlt_panel_tile<64>(sm, i0, j0, ld, Kb, db, kv);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Padded order <= 128 (one or two 64-blocks): the whole matrix in ONE CTA, three CTAs per SM.  Orders this small are
// bound by the latency of the diagonal-block chain times the number of matrices an SM holds at a time -- two with the
// per-column kernels (85 / 104 KB of shared memory per CTA).  Here the first diagonal block is factorised in tile R0
// with inv(L)' in its own upper triangle (ldlt_diag_factor<true>), block row 1 passes through tile R1 exactly as in the
// chain CTA of block column 0 (same products in the same order: the factors and the scratch triangle come out
// bit-identical to the per-column kernels), and its diagonal block is factorised in R0 again: 2 tiles + 706 doubles =
// 75.3 KB.  Measured (B = 4096): N = 64 0.233 -> 0.193 ms, N = 128 0.592 -> 0.49 ms -- 1.2x from 1.5x the matrices per
// SM: three chains on an SM slow each other.  Giving the pivot chain of co-resident CTAs different warps (sub-partitions)
// changed nothing (0.497 vs 0.491 ms); a 4-byte static __shared__ variable next to the dynamic tiles cost the per-column
// kernels 9 % (N = 768: 27.8 vs 25.45 ms) and this kernel 4 % -- the tiles want the dynamic segment at offset 0.
constexpr int SMALL_SMEM = (2 * NB * DP + DGM_MISC) * (int)sizeof(double);
static_assert(EP == DP, "the small kernel uses one pitch for the tiles of both roles");
static_assert(NB <= NB * WNPM, "1 / d of block 0 is parked in the (idle) -W panel");
__global__ void __launch_bounds__(256, 3) ldlt_small_kernel(int ld, const int32_t* __restrict__ Nvec, int Nfixed,
                                                             double* __restrict__ K, double* __restrict__ dvec,
                                                             int32_t* __restrict__ info, int32_t* __restrict__ nneg,
                                                             const int32_t* __restrict__ npos_expected, GfWork work) {
    const int b = gf_instance(work, blockIdx.x);
    if (b < 0) return;
    extern __shared__ double sm[];
    const int Np = padded_order(Nvec, Nfixed, b, ld);
    if (Np <= 0) {
        if (threadIdx.x == 0) { info[b] = 0; nneg[b] = 0; }
        return;
    }
    double* R0 = sm;
    double* R1 = sm + NB * DP;
    double* misc = sm + 2 * NB * DP;
    double* Kb = K + (size_t)b * ld * ld;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, g = lane >> 2, q = lane & 3;
    for (int e = tid; e < NB * NB / 2; e += 256) {
        const int r = e >> 5, c = (e & 31) * 2;
        const double2 v = *reinterpret_cast<const double2*>(Kb + (size_t)r * ld + c);
        R0[r * DP + c] = v.x;
        R0[r * DP + c + 1] = v.y;
    }
    __syncthreads();
    ldlt_diag_factor<true>(R0, nullptr, misc, b, ld, Nvec, Nfixed, 0, K, dvec, info, nneg, npos_expected);
    if (Np <= NB) return;
    // ---- block row 1 = the chain CTA of block column 0 (ldlt_chain_body with nothing to its left)
    const int i0 = NB;
    double* Xs = R0;          // storage layout of diagonal block 0: inv(L_00)' above the diagonal; later L[1,0]
    double* Cs = R1;          // A[1,0]; later W' = -C X'
    double* rinv = misc;      // 1 / d of block 0 (IEEE quotient, as the column kernels form it)
    for (int e = tid; e < NB * NB / 2; e += 256) {
        const int r = e >> 5, c = (e & 31) * 2;
        const double2 v = *reinterpret_cast<const double2*>(Kb + (size_t)(i0 + r) * ld + c);
        Cs[r * EP + c] = v.x;
        Cs[r * EP + c + 1] = v.y;
    }
    if (tid < NB) rinv[tid] = 1.0 / R0[tid * DP + tid];
    __syncthreads();
    {   // W = C X' (every warp 8 full rows), L = W D^{-1}
        double t[8][2];
#pragma unroll
        for (int ni = 0; ni < 8; ni++) { t[ni][0] = 0.0; t[ni][1] = 0.0; }
        const double* as = Cs + (wid * 8 + g) * EP + q;
#pragma unroll
        for (int kk = 0; kk < NB; kk += 4) {
            const double a = as[kk];
#pragma unroll
            for (int ni = 0; ni < 8; ni++) {
                if (kk < ni * 8 + 8) dmma884(t[ni][0], t[ni][1], a, xt_fragment(Xs, kk + q, ni * 8 + g));
            }
        }
        __syncthreads();  // all reads of Cs / Xs are done: overwrite them with W' and L
        const int r = wid * 8 + g;
#pragma unroll
        for (int ni = 0; ni < 8; ni++) {
            const int c = ni * 8 + 2 * q;
            const double r0 = rinv[c], r1 = rinv[c + 1];
            const double2 wv = make_double2(-t[ni][0], -t[ni][1]);
            const double2 lv = make_double2(t[ni][0] * r0, t[ni][1] * r1);
            *reinterpret_cast<double2*>(Cs + r * EP + c) = wv;
            *reinterpret_cast<double2*>(Xs + r * EP + c) = lv;
            *reinterpret_cast<double2*>(Kb + (size_t)(i0 + r) * ld + c) = lv;
            *reinterpret_cast<double2*>(Kb + (size_t)r * ld + i0 + c) = wv;
        }
    }
    __syncthreads();
    {   // S = A[1,1] + W' L'  (4 x 2 warps of 16 x 32, strictly upper quadrant skipped), accumulators from global
        const int wm2 = wid >> 1, wn2 = wid & 1;
        const bool on = !((wn2 == 1) && (wm2 < 2));
        double s2[2][4][2];
        if (on) {
#pragma unroll
            for (int mi = 0; mi < 2; mi++) {
                const int r = wm2 * 16 + mi * 8 + g;
#pragma unroll
                for (int ni = 0; ni < 4; ni++) {
                    const int c = wn2 * 32 + ni * 8 + 2 * q;
                    const double2 v = *reinterpret_cast<const double2*>(Kb + (size_t)(i0 + r) * ld + i0 + c);
                    s2[mi][ni][0] = v.x;
                    s2[mi][ni][1] = v.y;
                }
            }
            const double* as = Cs + (wm2 * 16 + g) * EP + q;
            const double* bs = Xs + (wn2 * 32 + g) * EP + q;
#pragma unroll 4
            for (int kk = 0; kk < NB; kk += 4) {
                double a[2], bf[4];
#pragma unroll
                for (int mi = 0; mi < 2; mi++) a[mi] = as[mi * 8 * EP + kk];
#pragma unroll
                for (int ni = 0; ni < 4; ni++) bf[ni] = bs[ni * 8 * EP + kk];
#pragma unroll
                for (int mi = 0; mi < 2; mi++)
#pragma unroll
                    for (int ni = 0; ni < 4; ni++) dmma884(s2[mi][ni][0], s2[mi][ni][1], a[mi], bf[ni]);
            }
        }
        __syncthreads();  // W' and L are no longer needed in shared memory: the diagonal tile goes into R0
        if (on) {
#pragma unroll
            for (int mi = 0; mi < 2; mi++) {
                const int r = wm2 * 16 + mi * 8 + g;
#pragma unroll
                for (int ni = 0; ni < 4; ni++) {
                    const int c = wn2 * 32 + ni * 8 + 2 * q;
                    R0[r * DP + c] = s2[mi][ni][0];
                    R0[r * DP + c + 1] = s2[mi][ni][1];
                }
            }
        }
    }
    __syncthreads();
    ldlt_diag_factor<true>(R0, nullptr, misc, b, ld, Nvec, Nfixed, 1, K, dvec, info, nneg, npos_expected);
}

// Panel tiles only, 64 rows each, three CTAs per SM (<= 85 registers, 74 KB shared memory): the rows below block
// k + 1 of block column k.  Launched beside the chain kernel of the same column.
constexpr int P64_SMEM = STAGES * 2 * NB * PSP * (int)sizeof(double);
__global__ void __launch_bounds__(256, 3) ldlt_panel64_kernel(int ld, const int32_t* __restrict__ Nvec, int Nfixed, int k,
                                                               double* __restrict__ K, const double* __restrict__ dvec,
                                                               GfWork work, int woff, KktSrc src) {
    const int b = gf_instance(work, woff + blockIdx.y);
    if (b < 0) return;
    extern __shared__ double sm[];
    const int Np = padded_order(Nvec, Nfixed, b, ld);
    const int j0 = k * NB;
    const int i0 = j0 + 2 * NB + blockIdx.x * NB;
    if (i0 >= Np) return;
    ldlt_panel_tile<64, 64>(sm, i0, j0, ld, K + (size_t)b * ld * ld, dvec + (size_t)b * ld, kkt_view(src, b));
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
            int i = lane;
            for (; i + 96 < j0; i += 128) {  // four loads in flight per lane
                const double r0 = __ldg(row + i), r1 = __ldg(row + i + 32), r2 = __ldg(row + i + 64), r3 = __ldg(row + i + 96);
                acc += r0 * v[i] + r1 * v[i + 32] + r2 * v[i + 64] + r3 * v[i + 96];
            }
            for (; i < j0; i += 32) acc += __ldg(row + i) * v[i];
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
            int ii = 0;
            for (; ii + 8 <= jb; ii += 8) {  // eight loads in flight per thread
                double h[8];
#pragma unroll
                for (int u = 0; u < 8; u++) h[u] = __ldg(Kb + (size_t)(j0 + ii + u) * ld + j);
#pragma unroll
                for (int u = 0; u < 8; u++) acc += h[u] * v[j0 + ii + u];
            }
            for (; ii < jb; ii++) acc += __ldg(Kb + (size_t)(j0 + ii) * ld + j) * v[j0 + ii];
            v[j] -= acc;
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < N; i += blockDim.x) rb[i] = v[i];
}

}  // namespace

// Two internal streams per device: a large batch is factorised in two halves whose block-column launches
// interleave, so the tail of one half's launch (few long chain CTAs left) is filled by the other half's CTAs.
namespace {
constexpr int LDLT_MAX_LANES = 4;
struct LdltLanes {
    cudaStream_t s[LDLT_MAX_LANES];    // chain (or whole-column) stream of each part of the batch
    cudaStream_t sp[LDLT_MAX_LANES];   // panel stream of each part (split mode)
    cudaEvent_t fork, join[LDLT_MAX_LANES], evc[LDLT_MAX_LANES], evp[LDLT_MAX_LANES];
    bool ok;
};
LdltLanes* ldlt_lanes() {
    static std::mutex mu;
    static LdltLanes lanes[64];
    static bool made[64] = {false};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    std::lock_guard<std::mutex> lock(mu);
    LdltLanes& L = lanes[dev];
    if (!made[dev]) {
        made[dev] = true;
        L.ok = true;
        for (int i = 0; i < LDLT_MAX_LANES; i++) {
            L.ok = L.ok && cudaStreamCreateWithFlags(&L.s[i], cudaStreamNonBlocking) == cudaSuccess;
            L.ok = L.ok && cudaStreamCreateWithFlags(&L.sp[i], cudaStreamNonBlocking) == cudaSuccess;
            L.ok = L.ok && cudaEventCreateWithFlags(&L.join[i], cudaEventDisableTiming) == cudaSuccess;
            L.ok = L.ok && cudaEventCreateWithFlags(&L.evc[i], cudaEventDisableTiming) == cudaSuccess;
            L.ok = L.ok && cudaEventCreateWithFlags(&L.evp[i], cudaEventDisableTiming) == cudaSuccess;
        }
        L.ok = L.ok && cudaEventCreateWithFlags(&L.fork, cudaEventDisableTiming) == cudaSuccess;
    }
    return L.ok ? &L : nullptr;
}
constexpr int LDLT_SPLIT_MIN = 1024;  // below this many matrices a launch has no tail worth hiding
}  // namespace

static int ldlt_factor_impl(int B, int ld, int Nmax, const int32_t* Nvec, double* K, double* dvec, int32_t* info,
                            int32_t* nneg, const int32_t* npos_expected, const int32_t* work,
                            const int32_t* nwork_dev, int nwork, void* stream, KktSrc src) {
    if (B <= 0 || ld <= 0 || (ld % NB) != 0 || Nmax < 0 || Nmax > ld || !K || !dvec || !info || !nneg)
        return GF_ERR_ARG;
    if (nwork <= 0 || Nmax == 0) return GF_OK;
    cudaStream_t s = (cudaStream_t)stream;
    GfWork w{work, nwork_dev};
    const int Np = ((Nmax + NB - 1) / NB) * NB;
    const int nblk = Np / NB;
#ifdef GF_LDLT_COL_SMEM_MIN  // developer experiment: force the occupancy of the column kernel
    constexpr int COL_SMEM = GF_LDLT_COL_SMEM_MIN;
#else
    constexpr int COL_SMEM = PN_SMEM;
#endif
    cudaFuncSetAttribute(ldlt_diag0_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DG_SMEM);
    cudaFuncSetAttribute(ldlt_column_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, COL_SMEM);
    // padded order <= 128: the whole matrix in one CTA, three CTAs per SM (GF_LDLT_SMALL=0: the per-column kernels)
    if (nblk <= 2 && src.H == nullptr) {
        const char* e = getenv("GF_LDLT_SMALL");
        if (e == nullptr || e[0] != '0') {
            cudaFuncSetAttribute(ldlt_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMALL_SMEM);
            ldlt_small_kernel<<<nwork, 256, SMALL_SMEM, s>>>(ld, Nvec, Nmax, K, dvec, info, nneg, npos_expected, w);
            return gf_launch_status();
        }
    }
    // GF_LDLT_WHOLE=1: one CTA per matrix (experiment, 5 % slower than the per-column launches; see ldlt_whole_kernel)
    static const bool whole_req = [] { const char* e = getenv("GF_LDLT_WHOLE"); return e != nullptr && e[0] == '1'; }();
    static_assert(DG_SMEM <= PN_SMEM, "the first diagonal block must fit into the column kernel's shared memory");
    if (whole_req && nblk > 1) {
        cudaFuncSetAttribute(ldlt_whole_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, COL_SMEM);
        ldlt_whole_kernel<<<nwork, 256, COL_SMEM, s>>>(ld, Nvec, Nmax, K, dvec, info, nneg, npos_expected, w, 0, src);
        return gf_launch_status();
    }
    // GF_LDLT_LANES: 0 = one stream, 2 (default) .. 4 = parts of the batch whose launches interleave
    static const int lanes_req = [] {
        const char* e = getenv("GF_LDLT_LANES");
        const int v = e == nullptr ? 2 : atoi(e);
        return v < 2 ? 1 : (v > LDLT_MAX_LANES ? LDLT_MAX_LANES : v);
    }();
    LdltLanes* L = (lanes_req > 1 && nwork >= LDLT_SPLIT_MIN && nblk > 2) ? ldlt_lanes() : nullptr;
    // The lane streams and their fork / join events are per device, shared by every caller: the whole fork -> launches ->
    // join sequence is issued under one lock, so two host threads (or two caller streams) factorising on the same device
    // cannot re-record each other's events between a record and the waits on it.  (A wait captures the event's state at
    // the time of the call, so holding the lock while ISSUING is sufficient; the kernels themselves overlap freely.)
    static std::mutex issue_mu;
    std::unique_lock<std::mutex> issue_lock(issue_mu, std::defer_lock);
    if (L != nullptr) issue_lock.lock();
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (L != nullptr && (cudaStreamIsCapturing(s, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone)) L = nullptr;
    static const bool split_on = [] { const char* e = getenv("GF_LDLT_P64"); return e != nullptr && e[0] == '1'; }();
    const int nlane = L != nullptr ? (split_on ? 2 : lanes_req) : 1;
    int off[LDLT_MAX_LANES], cnt[LDLT_MAX_LANES];
    cudaStream_t st[LDLT_MAX_LANES];
    for (int i = 0; i < nlane; i++) {
        off[i] = (int)((long)nwork * i / nlane);
        cnt[i] = (int)((long)nwork * (i + 1) / nlane) - off[i];
        st[i] = L != nullptr ? L->s[i] : s;
    }
    if (L != nullptr) {
        cudaEventRecord(L->fork, s);
        for (int i = 0; i < nlane; i++) cudaStreamWaitEvent(L->s[i], L->fork, 0);
    }
    // opt-in experiment (GF_LDLT_P64=1): +2 % on uniform batches, nothing on the ragged bench batch
    if (L != nullptr && split_on) {
        // Split mode: per half batch a chain stream (chain CTAs only, 2 per SM) and a panel stream (64-row panel tiles,
        // 3 per SM).  chain(k) needs panel(k-1); panel(k) needs chain(k-1) (the diagonal block of column k).
        cudaFuncSetAttribute(ldlt_panel64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, P64_SMEM);
        for (int i = 0; i < 2; i++) {
            ldlt_diag0_kernel<<<cnt[i], 256, DG_SMEM, st[i]>>>(ld, Nvec, Nmax, K, dvec, info, nneg, npos_expected, w, off[i], src);
            cudaEventRecord(L->evc[i], st[i]);
        }
        bool have_p[2] = {false, false};
        for (int k = 0; k + 1 < nblk; k++) {
            const int below = Np - (k + 2) * NB;
            const int t64 = below / NB;
            for (int i = 0; i < 2; i++) {
                if (have_p[i]) cudaStreamWaitEvent(st[i], L->evp[i], 0);
                if (t64 > 0) cudaStreamWaitEvent(L->sp[i], L->evc[i], 0);
                ldlt_column_kernel<<<cnt[i], 256, COL_SMEM, st[i]>>>(ld, Nvec, Nmax, k, K, dvec, info, nneg, npos_expected, w,
                                                                    off[i], cnt[i], 1, src, 0);
                if (t64 > 0)
                    ldlt_panel64_kernel<<<dim3(t64, cnt[i]), 256, P64_SMEM, L->sp[i]>>>(ld, Nvec, Nmax, k, K, dvec, w, off[i], src);
                cudaEventRecord(L->evc[i], st[i]);
                if (t64 > 0) {
                    cudaEventRecord(L->evp[i], L->sp[i]);
                    have_p[i] = true;
                }
            }
        }
        for (int i = 0; i < 2; i++)
            if (have_p[i]) cudaStreamWaitEvent(st[i], L->evp[i], 0);
    } else {
    static const int order_req = [] { const char* e = getenv("GF_LDLT_ORDER"); return e == nullptr ? 0 : atoi(e); }();
    // issue order interleaves the lanes launch by launch
    for (int i = 0; i < nlane; i++)
        ldlt_diag0_kernel<<<cnt[i], 256, DG_SMEM, st[i]>>>(ld, Nvec, Nmax, K, dvec, info, nneg, npos_expected, w, off[i], src);
    for (int k = 0; k + 1 < nblk; k++) {
        const int below = Np - (k + 2) * NB;              // rows under block k+1
        const int tiles = 1 + (below + TM - 1) / TM;      // chain CTA + 128-row panel tiles
        for (int i = 0; i < nlane; i++)
            ldlt_column_kernel<<<tiles * cnt[i], 256, COL_SMEM, st[i]>>>(ld, Nvec, Nmax, k, K, dvec, info, nneg,
                                                                         npos_expected, w, off[i], cnt[i], tiles, src,
                                                                         tiles > 1 ? order_req : 0);
    }
    }
    if (L != nullptr) {
        for (int i = 0; i < nlane; i++) {
            cudaEventRecord(L->join[i], L->s[i]);
            cudaStreamWaitEvent(s, L->join[i], 0);
        }
    }
    return gf_launch_status();
}

extern "C" int gf_ldlt_factor(int B, int ld, int Nmax, const int32_t* Nvec, double* K, double* dvec, int32_t* info,
                              int32_t* nneg, const int32_t* npos_expected, const int32_t* work,
                              const int32_t* nwork_dev, int nwork, void* stream) {
    KktSrc none;
    memset(&none, 0, sizeof(none));
    return ldlt_factor_impl(B, ld, Nmax, Nvec, K, dvec, info, nneg, npos_expected, work, nwork_dev, nwork, stream, none);
}

extern "C" int gf_kkt_ldlt_factor(int B, int n, int m, int ld, const double* H, const double* J, const int32_t* perm,
                                  const int32_t* nI, const double* dt, const double* rho, const int32_t* Nvec, double* K,
                                  double* dvec, int32_t* info, int32_t* nneg, const int32_t* work,
                                  const int32_t* nwork_dev, int nwork, void* stream) {
    if (n <= 0 || m < 0 || !H || !perm || !nI || !dt || !rho || !Nvec || (m > 0 && !J)) return GF_ERR_ARG;
    KktSrc src{H, J, perm, nI, dt, rho, n, m};
    return ldlt_factor_impl(B, ld, n + m, Nvec, K, dvec, info, nneg, nI, work, nwork_dev, nwork, stream, src);
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
