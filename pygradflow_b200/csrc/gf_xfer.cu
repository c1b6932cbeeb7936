// Host <-> device staging helpers of the host-buffer entry (pygradflow_b200/host_step.py).
//
// The Hessian of the Lagrangian is symmetric (pygradflow/problem.py:174-192 requires it; eval.py:183-211 validates
// it), so only its lower block triangle crosses PCIe: row block r (blk rows) is copied with the columns [0, end of
// its diagonal block) by one strided 3-D copy per row block, and the blocks above the diagonal are rebuilt on the
// device by a tiled transpose.  For n = 512, blk = 64 that is 56 % of the bytes of the full matrix.
#include <cstring>
#include "gf_common.cuh"
#include "../../include/gradflow_b200.h"

namespace {

// H[b][tj*32 + c][ti*32 + r] = H[b][ti*32 + r][tj*32 + c] for the 32 x 32 tiles (ti, tj) strictly below the
// diagonal blocks of size blk (those were not transferred in the upper triangle).
__global__ void symmetrize_kernel(int n, int blk, double* __restrict__ H) {
    __shared__ double tile[32][33];
    const int nt = (n + 31) / 32;
    // linear tile index -> (ti, tj), tj < ti
    int e = blockIdx.x, ti = 1;
    while (e >= ti) { e -= ti; ti++; }
    const int tj = e;
    if (ti >= nt) return;
    if ((ti * 32) / blk == (tj * 32) / blk) return;  // inside a diagonal block: transferred in full
    double* Hb = H + (size_t)blockIdx.y * n * n;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    for (int r = ty; r < 32; r += 8) {
        const int row = ti * 32 + r, col = tj * 32 + tx;
        tile[r][tx] = (row < n && col < n) ? Hb[(size_t)row * n + col] : 0.0;
    }
    __syncthreads();
    for (int c = ty; c < 32; c += 8) {
        const int row = tj * 32 + c, col = ti * 32 + tx;
        if (row < n && col < n) Hb[(size_t)row * n + col] = tile[tx][c];
    }
}

}  // namespace

extern "C" int gf_h2d_sym_lower(double* dst, const double* src_host, int cnt, int n, int blk, void* stream) {
    if (!dst || !src_host || cnt <= 0 || n <= 0 || blk <= 0) return GF_ERR_ARG;
    for (int r0 = 0; r0 < n; r0 += blk) {
        const int r1 = r0 + blk < n ? r0 + blk : n;
        cudaMemcpy3DParms p;
        memset(&p, 0, sizeof(p));
        p.srcPtr = make_cudaPitchedPtr((void*)(src_host + (size_t)r0 * n), (size_t)n * sizeof(double),
                                       (size_t)n * sizeof(double), (size_t)n);
        p.dstPtr = make_cudaPitchedPtr((void*)(dst + (size_t)r0 * n), (size_t)n * sizeof(double),
                                       (size_t)n * sizeof(double), (size_t)n);
        p.extent = make_cudaExtent((size_t)r1 * sizeof(double), (size_t)(r1 - r0), (size_t)cnt);
        p.kind = cudaMemcpyHostToDevice;
        const cudaError_t e = cudaMemcpy3DAsync(&p, (cudaStream_t)stream);
        if (e != cudaSuccess) return GF_ERR_CUDA(e);
    }
    return GF_OK;
}

extern "C" int gf_symmetrize_lower(double* H, int cnt, int n, int blk, void* stream) {
    if (!H || cnt <= 0 || n <= 0 || blk <= 0 || (blk % 32) != 0) return GF_ERR_ARG;
    const int nt = (n + 31) / 32;
    const int pairs = nt * (nt - 1) / 2;
    if (pairs == 0) return GF_OK;
    if (cnt > 65535) return GF_ERR_UNSUPPORTED;  // the instance index rides in grid.y
    symmetrize_kernel<<<dim3(pairs, cnt), 256, 0, (cudaStream_t)stream>>>(n, blk, H);
    return gf_launch_status();
}
