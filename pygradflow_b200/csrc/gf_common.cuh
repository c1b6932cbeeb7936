// Shared device helpers for libgradflow_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define GF_OK 0
#define GF_ERR_ARG (-1)
#define GF_ERR_UNSUPPORTED (-2)
#define GF_ERR_CUDA(e) (1000 + (int)(e))

#define GF_ACTIVE_SLACK 1e-8  // pygradflow/implicit_func.py:44

static inline int gf_launch_status() {
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? GF_OK : GF_ERR_CUDA(e);
}

// Work list: CTA w of the batch dimension processes instance work[w] (or w when work == NULL).
// nwork_dev (optional, device) lets the host launch an upper bound without reading the count back.
struct GfWork {
    const int32_t* list;
    const int32_t* count_dev;
    int off = 0;  // slot offset: a launch that covers the slots [off, off + grid) of the list (batches split over streams)
};

__device__ __forceinline__ int gf_instance(const GfWork& w, int cta) {
    cta += w.off;
    if (w.count_dev != nullptr && cta >= *w.count_dev) return -1;
    return w.list != nullptr ? w.list[cta] : cta;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Deterministic block reductions (fixed tree: lanes by xor-shuffle, then warp 0 over the warp partials).
// `scratch` must hold >= 32 doubles.  All threads of the block must call.
__device__ __forceinline__ double block_sum(double v, double* scratch) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) scratch[wid] = v;
    __syncthreads();
    double r = (threadIdx.x < nw) ? scratch[threadIdx.x] : 0.0;
    if (wid == 0) r = warp_sum(r);
    if (threadIdx.x == 0) scratch[0] = r;
    __syncthreads();
    return scratch[0];
}

__device__ __forceinline__ double block_max(double v, double* scratch) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_max(v);
    __syncthreads();
    if (lane == 0) scratch[wid] = v;
    __syncthreads();
    double r = (threadIdx.x < nw) ? scratch[threadIdx.x] : -INFINITY;
    if (wid == 0) r = warp_max(r);
    if (threadIdx.x == 0) scratch[0] = r;
    __syncthreads();
    return scratch[0];
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// D(8x8) += A(8x4, row) * B(4x8, col) on the FP64 tensor pipe (SASS: DMMA.8x8x4).
// Fragment layout (lane t): a = A[t/4][t%4], b = B[t%4][t/4], c0/c1 = C[t/4][2*(t%4) + {0,1}].
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}
