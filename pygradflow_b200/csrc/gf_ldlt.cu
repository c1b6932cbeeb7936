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
// 64-block.  d is kept on the diagonal of K and, contiguously, in dvec[b].  The strict upper triangle of
// every 64x64 DIAGONAL block is scratch: it receives inv(L_kk)' for the panel kernel.
//
// Left-looking by 64-wide block columns, two batched launches per block column k (j0 = 64 k):
//   diag(k)   C = A[k,k] - sum_{p<k} (L[k,p] D_p) L[k,p]'  (DMMA, cp.async pipeline), then in shared memory
//             C = L_kk D_k L_kk' and X = inv(L_kk) (blocked 16/32/64), one CTA per matrix;
//   panel(k)  128x64 tiles below the diagonal: C = A[i,k] - sum_{p<k} (L[i,p] D_p) L[k,p]'  (DMMA),
//             then the triangular solve as one more DMMA product L[i,k] = (C X') D_k^{-1}.
// Every matrix entry is read and written once, operands stream once per block column
// (~ 8 N^3 / (6*64) bytes per matrix), so the factorisation is bound by the FP64 pipe, not by HBM.
#include "gf_common.cuh"
#include "../../include/gradflow_b200.h"

namespace {

constexpr int NB = 64;
constexpr int KC = 16;       // k-chunk per pipeline stage
constexpr int SP = KC + 4;   // smem row pitch of a stage (bank-conflict-free DMMA fragment loads)
constexpr int EP = NB + 4;   // smem row pitch of 64-wide tiles in the epilogues (same property)

__device__ __forceinline__ int padded_order(const int32_t* Nvec, int Nfixed, int b, int ld) {
    const int N = Nvec != nullptr ? Nvec[b] : Nfixed;
    const int Np = ((N + NB - 1) / NB) * NB;
    return Np < ld ? Np : ld;
}

// dst (M x N, pitch ldd) = -A (M x Kd, pitch lda) * Bm (Kd x N, pitch ldb), all in shared memory.
__device__ __forceinline__ void smem_neg_matmul(double* dst, int ldd, const double* A, int lda, const double* Bm,
                                                int ldb, int M, int N, int Kd) {
    for (int e = threadIdx.x; e < M * N; e += blockDim.x) {
        const int r = e / N, c = e - r * N;
        double acc = 0.0;
#pragma unroll 8
        for (int p = 0; p < Kd; p++) acc = fma(A[r * lda + p], Bm[p * ldb + c], acc);
        dst[r * ldd + c] = -acc;
    }
}

// ------------------------------------------------------------------------------------------------
constexpr int DG_STAGES = 3;
constexpr int DG_BW = 4;  // panel width of the in-shared-memory factorisation
constexpr int XP = NB + 1;  // pitch of the inverse in shared memory
constexpr int DG_OPS = (DG_STAGES * NB * SP > NB * XP) ? DG_STAGES * NB * SP : NB * XP;
constexpr int DG_SMEM = (NB * (NB + 1) + DG_OPS + NB + 32 * 33 + NB * (DG_BW + 1) + 2) * (int)sizeof(double);

__device__ __forceinline__ void ldlt_diag_body(double* sm, int b, int ld, const int32_t* __restrict__ Nvec,
                                               int Nfixed, int k, double* __restrict__ K,
                                               double* __restrict__ dvec, int32_t* __restrict__ info,
                                               int32_t* __restrict__ nneg,
                                               const int32_t* __restrict__ npos_expected) {
    const int Np = padded_order(Nvec, Nfixed, b, ld);
    const int j0 = k * NB;
    if (j0 >= Np) return;
    double(*S)[NB + 1] = reinterpret_cast<double(*)[NB + 1]>(sm);   // 64 x 65
    double* Ops = sm + NB * (NB + 1);                                 // DG_STAGES x 64 x SP, later X = inv(L)
    double* rinv = Ops + DG_OPS;                                      // 64
    double* T = rinv + NB;                                            // 32 x 33
    double* Wp = T + 32 * 33;                                         // 64 x (DG_BW + 1): W = L D of the current panel
    int* flags = reinterpret_cast<int*>(Wp + NB * (DG_BW + 1));
    int& s_bad = flags[0];
    int& s_neg = flags[1];
    int& s_sign = flags[2];
    double* Kb = K + (size_t)b * ld * ld;
    const double* db = dvec + (size_t)b * ld;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) { s_bad = 0x7fffffff; s_neg = 0; s_sign = 0; }

    // ---- phase A: diagonal tile minus the contributions of the block columns to its left (DMMA)
    {
        const int wm = wid >> 1, wn = wid & 1;  // 4 x 2 warps, 16 x 32 each
        const int g = lane >> 2, q = lane & 3;
        const bool upper_quadrant = (wn == 1) && (wm < 2);  // rows 0..31 x cols 32..63: never read
        double acc[2][4][2];
#pragma unroll
        for (int mi = 0; mi < 2; mi++) {
            const int row = j0 + wm * 16 + mi * 8 + g;
#pragma unroll
            for (int ni = 0; ni < 4; ni++) {
                const int col = j0 + wn * 32 + ni * 8 + 2 * q;
                const double2 v = *reinterpret_cast<const double2*>(Kb + (size_t)row * ld + col);
                acc[mi][ni][0] = v.x;
                acc[mi][ni][1] = v.y;
            }
        }
        const int nchunks = j0 / KC;
        auto load_stage = [&](int chunk, int stage) {
            const int p0 = chunk * KC;
            double* os = Ops + stage * NB * SP;
#pragma unroll
            for (int t = 0; t < 2; t++) {
                const int piece = tid + t * 256;  // 512 pieces of 16 B
                const int r = piece >> 3, part = piece & 7;
                cp_async16(os + r * SP + part * 2, Kb + (size_t)(j0 + r) * ld + p0 + part * 2);
            }
        };
#pragma unroll
        for (int s = 0; s < DG_STAGES - 1; s++) {
            if (s < nchunks) load_stage(s, s);
            cp_async_commit();
        }
        for (int c = 0; c < nchunks; c++) {
            cp_async_wait<DG_STAGES - 2>();
            __syncthreads();
            const int nxt = c + DG_STAGES - 1;
            if (nxt < nchunks) load_stage(nxt, nxt % DG_STAGES);
            cp_async_commit();
            const double* os = Ops + (c % DG_STAGES) * NB * SP;
            const double* as = os + (wm * 16 + g) * SP + q;
            const double* bs = os + (wn * 32 + g) * SP + q;
            const double* dp = db + c * KC + q;
#pragma unroll
            for (int kk = 0; kk < KC; kk += 4) {
                const double nd = -__ldg(dp + kk);
                double a[2], bf[4];
#pragma unroll
                for (int mi = 0; mi < 2; mi++) a[mi] = as[mi * 8 * SP + kk] * nd;
#pragma unroll
                for (int ni = 0; ni < 4; ni++) bf[ni] = bs[ni * 8 * SP + kk];
                if (!upper_quadrant) {
#pragma unroll
                    for (int mi = 0; mi < 2; mi++)
#pragma unroll
                        for (int ni = 0; ni < 4; ni++) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], bf[ni]);
                }
            }
        }
        cp_async_wait<0>();
#pragma unroll
        for (int mi = 0; mi < 2; mi++) {
            const int r = wm * 16 + mi * 8 + g;
#pragma unroll
            for (int ni = 0; ni < 4; ni++) {
                const int c = wn * 32 + ni * 8 + 2 * q;
                S[r][c] = acc[mi][ni][0];
                S[r][c + 1] = acc[mi][ni][1];
            }
        }
    }
    __syncthreads();

    // ---- phase B: C = L D L' in place, right-looking in panels of DG_BW columns on W = L D (scaled to L
    // afterwards).  Every thread factors the small diagonal block redundantly in registers, so a panel
    // costs two barriers instead of DG_BW.
    {
        const int i = tid & 63, cg = tid >> 6;
        constexpr int BW = DG_BW;
        for (int jb = 0; jb < NB; jb += BW) {
            double Dg[BW][BW], rj[BW];
#pragma unroll
            for (int r = 0; r < BW; r++)
#pragma unroll
                for (int c = 0; c <= r; c++) Dg[r][c] = S[jb + r][jb + c];
#pragma unroll
            for (int j = 0; j < BW; j++) {
                const double d = Dg[j][j];
                rj[j] = (d != 0.0) ? __drcp_rn(d) : 0.0;
#pragma unroll
                for (int i2 = j + 1; i2 < BW; i2++) {
                    const double l = Dg[i2][j] * rj[j];
#pragma unroll
                    for (int c = j + 1; c <= i2; c++) Dg[i2][c] = fma(-l, Dg[c][j], Dg[i2][c]);
                }
            }
            double l[BW], w[BW];
            const bool below = i >= jb + BW;
            if (below) {
#pragma unroll
                for (int c = 0; c < BW; c++) w[c] = S[i][jb + c];
#pragma unroll
                for (int c = 0; c < BW; c++) {
#pragma unroll
                    for (int p = 0; p < c; p++) w[c] = fma(-l[p], Dg[c][p], w[c]);
                    l[c] = w[c] * rj[c];
                }
                if (cg == 0) {
#pragma unroll
                    for (int c = 0; c < BW; c++) Wp[i * (BW + 1) + c] = w[c];
                }
            }
            __syncthreads();  // Wp complete; nobody reads S[., jb..jb+7] below this line any more
            if (below) {
                if (cg == 0) {
#pragma unroll
                    for (int c = 0; c < BW; c++) S[i][jb + c] = w[c];
                }
                const int cbeg = jb + BW;
                const int cstart = cbeg + ((cg - cbeg) & 3);
#pragma unroll 4
                for (int c = cstart; c <= i; c += 4) {
                    double acc = S[i][c];
#pragma unroll
                    for (int p = 0; p < BW; p++) acc = fma(-l[p], Wp[c * (BW + 1) + p], acc);
                    S[i][c] = acc;
                }
            } else if (i >= jb && cg == 0) {
                const int r = i - jb;
#pragma unroll
                for (int rr = 0; rr < BW; rr++)
                    if (rr == r) {
#pragma unroll
                        for (int c = 0; c <= rr; c++) S[i][jb + c] = Dg[rr][c];
                    }
            }
            __syncthreads();
        }
    }
    if (tid < NB) {
        const double d = S[tid][tid];
        if (!(isfinite(d)) || d == 0.0) atomicMin(&s_bad, tid + 1);
        const int N = Nvec != nullptr ? Nvec[b] : Nfixed;
        const int j = j0 + tid;
        if (d < 0.0 && j < N) atomicAdd(&s_neg, 1);
        // quasi-definite sign pattern: the first npos pivots positive, the remaining (up to N) negative
        if (npos_expected != nullptr && j < N && ((j < npos_expected[b]) != (d > 0.0))) s_sign = 1;
        dvec[(size_t)b * ld + j] = d;
        rinv[tid] = 1.0 / d;
    }
    __syncthreads();
    for (int e = tid; e < NB * NB; e += blockDim.x) {
        const int r = e >> 6, c = e & 63;
        if (c < r) S[r][c] *= rinv[c];  // W -> L
    }
    __syncthreads();

    // ---- phase C: X = inv(L) (unit lower), blocked 16 -> 32 -> 64; X and the scratch T live in Ops
    double* X = Ops;                    // 64 x XP
    for (int e = tid; e < NB * NB; e += blockDim.x) {
        const int r = e >> 6, c = e & 63;
        X[r * XP + c] = (r == c) ? 1.0 : 0.0;
    }
    __syncthreads();
    if (tid < NB) {  // the four 16x16 diagonal blocks, one thread per column
        const int blk = tid >> 4, j = tid & 15, o = blk * 16;
        double x[16];
#pragma unroll
        for (int i = 0; i < 16; i++) x[i] = (i == j) ? 1.0 : 0.0;
#pragma unroll
        for (int i = 1; i < 16; i++) {
            double s = 0.0;
#pragma unroll
            for (int p = 0; p < i; p++) s = fma(S[o + i][o + p], x[p], s);
            if (i > j) x[i] = -s;
        }
#pragma unroll
        for (int i = 0; i < 16; i++)
            if (i > j) X[(o + i) * XP + o + j] = x[i];
    }
    __syncthreads();
    // 32-level: X[h+16:h+32, h:h+16] = -X11' * (L10 * X00), h in {0, 32}
    for (int h = 0; h < NB; h += 32) {
        for (int e = tid; e < 16 * 16; e += blockDim.x) {  // T = L10 * X00
            const int r = e >> 4, c = e & 15;
            double acc = 0.0;
#pragma unroll
            for (int p = 0; p < 16; p++) acc = fma(S[h + 16 + r][h + p], X[(h + p) * XP + h + c], acc);
            T[r * 33 + c] = acc;
        }
        __syncthreads();
        smem_neg_matmul(X + (h + 16) * XP + h, XP, X + (h + 16) * XP + h + 16, XP, T, 33, 16, 16, 16);
        __syncthreads();
    }
    // 64-level: X[32:64, 0:32] = -X[32:64, 32:64] * (L[32:64, 0:32] * X[0:32, 0:32])
    for (int e = tid; e < 32 * 32; e += blockDim.x) {
        const int r = e >> 5, c = e & 31;
        double acc = 0.0;
#pragma unroll 8
        for (int p = 0; p < 32; p++) acc = fma(S[32 + r][p], X[p * XP + c], acc);
        T[r * 33 + c] = acc;
    }
    __syncthreads();
    smem_neg_matmul(X + 32 * XP, XP, X + 32 * XP + 32, XP, T, 33, 32, 32, 32);
    __syncthreads();

    // ---- write back: L (strict lower), d (diagonal), inv(L)' (strict upper)
    for (int e = tid; e < NB * NB; e += blockDim.x) {
        const int r = e >> 6, c = e & 63;
        double v;
        if (c < r) v = S[r][c];
        else if (c == r) v = S[r][r];
        else v = X[c * XP + r];  // K[j0 + r][j0 + c] = X[c][r], c > r
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
}

// ------------------------------------------------------------------------------------------------
constexpr int TM = 128, TN = 64, STAGES = 2;
constexpr int PKC = 32;        // k-chunk per pipeline stage of the panel kernel
constexpr int PSP = PKC + 4;   // its smem row pitch (== 4 mod 16: conflict-free fragment loads)
constexpr int PN_SMEM_PIPE = (STAGES * (TM + TN) * PSP + STAGES * PKC) * (int)sizeof(double);
constexpr int PN_SMEM_EPI = (TM + TN) * EP * (int)sizeof(double);
constexpr int PN_SMEM = PN_SMEM_PIPE > PN_SMEM_EPI ? PN_SMEM_PIPE : PN_SMEM_EPI;

// One ROWS x 64 tile of block column k (rows i0.., all inside the padded order): update + triangular solve.
// ROWS = 128: 4 x 2 warps of 32 x 32;  ROWS = 64 (odd remainder block): 2 x 4 warps of 32 x 16.
template <int ROWS>
__device__ __forceinline__ void ldlt_panel_tile(double* sm, int i0, int j0, int ld, double* __restrict__ Kb,
                                                const double* __restrict__ db) {
    constexpr int WN = (ROWS == 128) ? 2 : 4;   // warps along the 64 columns
    constexpr int NI = 8 / WN;                  // 8-column DMMA tiles per warp
    constexpr int WC = NI * 8;                  // columns per warp
    constexpr int AT = ROWS / 16;               // 16-byte pieces per thread per stage for the A rows
    double* As = sm;                            // STAGES x TM x PSP (ROWS rows used)
    double* Bs = sm + STAGES * TM * PSP;        // STAGES x TN x PSP
    double* Ds = Bs + STAGES * TN * PSP;        // STAGES x PKC : d of the chunk
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int wm = wid / WN, wn = wid % WN;
    const int g = lane >> 2, q = lane & 3;
    const int nchunks = j0 / PKC;

    // Each thread copies one 16-byte piece of AT A rows and 4 B rows per stage: row = (tid >> 4) + 16 t,
    // piece = tid & 15; pointers advance by constants.  Threads 0..15 also copy the chunk's 32 pivots.
    const int lrow = tid >> 4, lpart = (tid & 15) * 2;
    const double* gA = Kb + (size_t)(i0 + lrow) * ld + lpart;
    const double* gB = Kb + (size_t)(j0 + lrow) * ld + lpart;
    const size_t gstride = (size_t)16 * ld;
    double* sA = As + lrow * PSP + lpart;
    double* sB = Bs + lrow * PSP + lpart;
    auto load_stage = [&](int chunk, int stage) {
        const double* ga = gA + chunk * PKC;
        const double* gb = gB + chunk * PKC;
        double* sa = sA + stage * TM * PSP;
        double* sb = sB + stage * TN * PSP;
#pragma unroll
        for (int t = 0; t < AT; t++) cp_async16(sa + t * 16 * PSP, ga + t * gstride);
#pragma unroll
        for (int t = 0; t < 4; t++) cp_async16(sb + t * 16 * PSP, gb + t * gstride);
        if (tid < PKC / 2) cp_async16(Ds + stage * PKC + tid * 2, db + chunk * PKC + tid * 2);
    };
#pragma unroll
    for (int s = 0; s < STAGES - 1; s++) {
        if (s < nchunks) load_stage(s, s);
        cp_async_commit();
    }
    // accumulators start from the current A tile (these loads overlap the first pipeline stage)
    double acc[4][NI][2];
#pragma unroll
    for (int mi = 0; mi < 4; mi++) {
        const double* rowp = Kb + (size_t)(i0 + wm * 32 + mi * 8 + g) * ld + j0 + wn * WC + 2 * q;
#pragma unroll
        for (int ni = 0; ni < NI; ni++) {
            const double2 v = *reinterpret_cast<const double2*>(rowp + ni * 8);
            acc[mi][ni][0] = v.x;
            acc[mi][ni][1] = v.y;
        }
    }
    for (int c = 0; c < nchunks; c++) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        const int nxt = c + STAGES - 1;
        if (nxt < nchunks) load_stage(nxt, nxt % STAGES);
        cp_async_commit();
        const double* as = As + (c % STAGES) * TM * PSP + (wm * 32 + g) * PSP + q;
        const double* bs = Bs + (c % STAGES) * TN * PSP + (wn * WC + g) * PSP + q;
        const double* ds = Ds + (c % STAGES) * PKC + q;
        // software pipeline: the fragments of k-step kk+4 are fetched while the DMMAs of kk issue
        double a[4], bf[NI], dcur;
#pragma unroll
        for (int mi = 0; mi < 4; mi++) a[mi] = as[mi * 8 * PSP];
#pragma unroll
        for (int ni = 0; ni < NI; ni++) bf[ni] = bs[ni * 8 * PSP];
        dcur = ds[0];
#pragma unroll
        for (int kk = 0; kk < PKC; kk += 4) {
            double an[4], bn[NI], dn = 0.0;
            if (kk + 4 < PKC) {
#pragma unroll
                for (int mi = 0; mi < 4; mi++) an[mi] = as[mi * 8 * PSP + kk + 4];
#pragma unroll
                for (int ni = 0; ni < NI; ni++) bn[ni] = bs[ni * 8 * PSP + kk + 4];
                dn = ds[kk + 4];
            }
#pragma unroll
            for (int mi = 0; mi < 4; mi++) a[mi] *= -dcur;
#pragma unroll
            for (int mi = 0; mi < 4; mi++)
#pragma unroll
                for (int ni = 0; ni < NI; ni++) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], bf[ni]);
            if (kk + 4 < PKC) {
#pragma unroll
                for (int mi = 0; mi < 4; mi++) a[mi] = an[mi];
#pragma unroll
                for (int ni = 0; ni < NI; ni++) bf[ni] = bn[ni];
                dcur = dn;
            }
        }
    }
    cp_async_wait<0>();
    __syncthreads();  // every warp is done with the pipeline buffers

    // ---- epilogue: L[i,k] = (C X') D^{-1} with X = inv(L_kk) read from the diagonal block's upper triangle
    double* Cs = sm;                 // ROWS x EP
    double* Xs = sm + TM * EP;       // TN x EP : Xs[n][kk] = X[n][kk]
#pragma unroll
    for (int mi = 0; mi < 4; mi++) {
        const int r = wm * 32 + mi * 8 + g;
#pragma unroll
        for (int ni = 0; ni < NI; ni++) {
            const int c = wn * WC + ni * 8 + 2 * q;
            *reinterpret_cast<double2*>(Cs + r * EP + c) = make_double2(acc[mi][ni][0], acc[mi][ni][1]);
        }
    }
    for (int e = tid; e < NB * NB; e += 256) {
        const int kk = e >> 6, n = e & 63;  // coalesced read of storage row j0 + kk
        double v = 0.0;
        if (n > kk) v = Kb[(size_t)(j0 + kk) * ld + j0 + n];
        else if (n == kk) v = 1.0;
        Xs[n * EP + kk] = v;
    }
    __syncthreads();
#pragma unroll
    for (int mi = 0; mi < 4; mi++)
#pragma unroll
        for (int ni = 0; ni < NI; ni++) { acc[mi][ni][0] = 0.0; acc[mi][ni][1] = 0.0; }
    {
        const double* as = Cs + (wm * 32 + g) * EP + q;
        const double* bs = Xs + (wn * WC + g) * EP + q;
        // X[n][kk] = 0 for kk > n: the 8 columns starting at n0 only need kk < n0 + 8
        const int kend = wn * WC + WC;
        for (int kk = 0; kk < kend; kk += 4) {
            double a[4];
#pragma unroll
            for (int mi = 0; mi < 4; mi++) a[mi] = as[mi * 8 * EP + kk];
#pragma unroll
            for (int ni = 0; ni < NI; ni++) {
                if (kk < wn * WC + ni * 8 + 8) {
                    const double bf = bs[ni * 8 * EP + kk];
#pragma unroll
                    for (int mi = 0; mi < 4; mi++) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], bf);
                }
            }
        }
    }
#pragma unroll
    for (int ni = 0; ni < NI; ni++) {
        const int col = j0 + wn * WC + ni * 8 + 2 * q;
        const double r0 = 1.0 / __ldg(db + col), r1 = 1.0 / __ldg(db + col + 1);
#pragma unroll
        for (int mi = 0; mi < 4; mi++) {
            const int row = i0 + wm * 32 + mi * 8 + g;
            *reinterpret_cast<double2*>(Kb + (size_t)row * ld + col) =
                make_double2(acc[mi][ni][0] * r0, acc[mi][ni][1] * r1);
        }
    }
}

__device__ __forceinline__ void ldlt_panel_body(double* sm, int b, int tile, int ld,
                                                const int32_t* __restrict__ Nvec, int Nfixed, int k,
                                                double* __restrict__ K, const double* __restrict__ dvec) {
    const int Np = padded_order(Nvec, Nfixed, b, ld);
    const int j0 = k * NB;
    const int i0 = j0 + NB + tile * TM;
    if (i0 >= Np) return;
    double* Kb = K + (size_t)b * ld * ld;
    const double* db = dvec + (size_t)b * ld;
    if (Np - i0 >= TM) ldlt_panel_tile<128>(sm, i0, j0, ld, Kb, db);
    else ldlt_panel_tile<64>(sm, i0, j0, ld, Kb, db);   // odd remainder block: Np - i0 == 64
}

// ------------------------------------------------------------------------------------------------
// Launch wrappers.  `fused` interleaves, per matrix, the diag role of block column k+1 (blockIdx.x == 0) with
// the panel tiles 1.. of block column k: the latency-bound factor / inverse phases of the diagonal block then
// share an SM with a DMMA-bound panel CTA instead of idling the GPU between launches (look-ahead).
__global__ void __launch_bounds__(256, 3) ldlt_diag_kernel(int ld, const int32_t* __restrict__ Nvec, int Nfixed, int k,
                                                        double* __restrict__ K, double* __restrict__ dvec,
                                                        int32_t* __restrict__ info, int32_t* __restrict__ nneg,
                                                        const int32_t* __restrict__ npos_expected, GfWork work) {
    const int b = gf_instance(work, blockIdx.x);
    if (b < 0) return;
    extern __shared__ double sm[];
    ldlt_diag_body(sm, b, ld, Nvec, Nfixed, k, K, dvec, info, nneg, npos_expected);
}

__global__ void __launch_bounds__(256, 2) ldlt_panel_kernel(int ld, const int32_t* __restrict__ Nvec, int Nfixed,
                                                            int k, int tile0, double* __restrict__ K,
                                                            const double* __restrict__ dvec, GfWork work) {
    const int b = gf_instance(work, blockIdx.y);
    if (b < 0) return;
    extern __shared__ double sm[];
    ldlt_panel_body(sm, b, tile0 + blockIdx.x, ld, Nvec, Nfixed, k, K, dvec);
}

__global__ void __launch_bounds__(256, 2) ldlt_fused_kernel(int ld, const int32_t* __restrict__ Nvec, int Nfixed,
                                                            int k, double* __restrict__ K, double* __restrict__ dvec,
                                                            int32_t* __restrict__ info, int32_t* __restrict__ nneg,
                                                            const int32_t* __restrict__ npos_expected, GfWork work) {
    const int b = gf_instance(work, blockIdx.y);
    if (b < 0) return;
    extern __shared__ double sm[];
    if (blockIdx.x == 0) ldlt_diag_body(sm, b, ld, Nvec, Nfixed, k + 1, K, dvec, info, nneg, npos_expected);
    else ldlt_panel_body(sm, b, blockIdx.x, ld, Nvec, Nfixed, k, K, dvec);
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
                              int32_t* nneg, const int32_t* npos_expected, const int32_t* work,
                              const int32_t* nwork_dev, int nwork, void* stream) {
    if (B <= 0 || ld <= 0 || (ld % NB) != 0 || Nmax < 0 || Nmax > ld || !K || !dvec || !info || !nneg)
        return GF_ERR_ARG;
    if (nwork <= 0 || Nmax == 0) return GF_OK;
    cudaStream_t s = (cudaStream_t)stream;
    GfWork w{work, nwork_dev};
    const int Np = ((Nmax + NB - 1) / NB) * NB;
    const int nblk = Np / NB;
    constexpr int FUSED_SMEM = PN_SMEM > DG_SMEM ? PN_SMEM : DG_SMEM;
    cudaFuncSetAttribute(ldlt_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DG_SMEM);
    cudaFuncSetAttribute(ldlt_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PN_SMEM);
    cudaFuncSetAttribute(ldlt_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FUSED_SMEM);
    ldlt_diag_kernel<<<nwork, 256, DG_SMEM, s>>>(ld, Nvec, Nmax, 0, K, dvec, info, nneg, npos_expected, w);
    for (int k = 0; k + 1 < nblk; k++) {
        const int j0 = k * NB;
        const int tiles = (Np - j0 - NB + TM - 1) / TM;  // >= 1
        // tile 0 holds the rows diag(k+1) needs; everything else of panel(k) runs beside diag(k+1)
        ldlt_panel_kernel<<<dim3(1, nwork), 256, PN_SMEM, s>>>(ld, Nvec, Nmax, k, 0, K, dvec, w);
        if (tiles > 1)
            ldlt_fused_kernel<<<dim3(tiles, nwork), 256, FUSED_SMEM, s>>>(ld, Nvec, Nmax, k, K, dvec, info, nneg,
                                                                          npos_expected, w);
        else  // nothing to interleave with: the standalone diag kernel fits three CTAs per SM
            ldlt_diag_kernel<<<nwork, 256, DG_SMEM, s>>>(ld, Nvec, Nmax, k + 1, K, dvec, info, nneg, npos_expected,
                                                         w);
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
