// Power-of-two scaling of the problem callbacks (pygradflow/scale.py:153-231, ScaledProblem): every scaled quantity is
// ldexp(original, integer exponent), exact in FP64.  One elementwise kernel for all of them:
//   out[b][r][c] = ldexp(in[b][r][c], sr * rw[b][r] + sc * cw[b][c] + so * ow[b])
// (vectors are matrices with one row).  HBM-bound: 16 bytes per element, rows stream coalesced; in == out allowed.
#include "gf_common.cuh"
#include "../../include/gradflow_b200.h"

namespace {

__global__ void ldexp_kernel(int rows, int cols, const double* __restrict__ in, const int32_t* __restrict__ rw, int sr,
                             const int32_t* __restrict__ cw, int sc, const int32_t* __restrict__ ow, int so,
                             double* __restrict__ out, GfWork work) {
    const int b = gf_instance(work, blockIdx.y);
    if (b < 0) return;
    const size_t base = (size_t)b * rows * cols;
    const int e0 = so != 0 && ow != nullptr ? so * ow[b] : 0;
    const int total = rows * cols;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int r = idx / cols, c = idx - r * cols;
        int e = e0;
        if (rw != nullptr) e += sr * rw[(size_t)b * rows + r];
        if (cw != nullptr) e += sc * cw[(size_t)b * cols + c];
        out[base + idx] = ldexp(in[base + idx], e);
    }
}

}  // namespace

extern "C" int gf_ldexp(int B, int rows, int cols, const double* in, const int32_t* rw, int sr, const int32_t* cw,
                        int sc, const int32_t* ow, int so, double* out, const int32_t* work,
                        const int32_t* nwork_dev, int nwork, void* stream) {
    if (B <= 0 || rows < 0 || cols < 0 || !in || !out) return GF_ERR_ARG;
    if (nwork <= 0 || rows == 0 || cols == 0) return GF_OK;
    const long total = (long)rows * cols;
    if (total > (1L << 30) || nwork > 65535) return GF_ERR_UNSUPPORTED;  // the instance index rides in grid.y
    int gx = (int)((total + 1023) / 1024);
    if (gx > 64) gx = 64;
    ldexp_kernel<<<dim3(gx, nwork), 256, 0, (cudaStream_t)stream>>>(rows, cols, in, rw, sr, cw, sc, ow, so, out,
                                                                   GfWork{work, nwork_dev});
    return gf_launch_status();
}
