// Microbenchmark: does a cheap "yield" in a DMMA-streaming loop give a latency-bound chain on the same sub-partition
// fair access to the FP64 pipe, and what does it cost the stream?  Probe: dependent pivot step (shuffle + MUFU + 3 DFMA)
// in warps 0..3; streamers: 2 warps per sub-partition, 16 independent DMMAs per group, then YIELD.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ long long g_res[4];
template <int YIELD, int GROUPS>
__global__ void bench(double* out, int iters, int probe_on) {
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __shared__ volatile int stop;
    __shared__ double sdat[64];
    if (threadIdx.x == 0) stop = 0;
    if (threadIdx.x < 64) sdat[threadIdx.x] = 1.0 + threadIdx.x * 1e-9;
    __syncthreads();
    if (wid >= 4) {
        double c[16][2];
        double a = 1.0 + lane * 1e-9, b = 0.5;
#pragma unroll
        for (int i = 0; i < 16; i++) { c[i][0] = i; c[i][1] = -i; }
        long long n = 0;
        const long long t0 = clock64();
        int x = lane;
        while (!stop) {
#pragma unroll
            for (int r = 0; r < GROUPS; r++) {
#pragma unroll
                for (int i = 0; i < 16; i++) dmma884(c[i][0], c[i][1], a, b);
                if (YIELD == 1) asm volatile("nanosleep.u32 0;");
                if (YIELD == 2) { asm volatile("{ .reg .b32 t; add.u32 t, %0, 1; mul.lo.u32 t, t, 3; add.u32 %0, t, 7; }" : "+r"(x)); if (x == 123456789) stop = 2; }
                if (YIELD == 3) __syncwarp();
                if (YIELD == 4) { a = sdat[(lane + (int)n) & 63]; }   // a dependent shared-memory load feeding the next group
                if (YIELD == 5) asm volatile("nanosleep.u32 20;");
            }
            n += GROUPS;
        }
        const long long t1 = clock64();
        if (wid == 4 && lane == 0) { g_res[1] = n * 16; g_res[2] = t1 - t0; }
        double s = 0;
#pragma unroll
        for (int i = 0; i < 16; i++) s += c[i][0] + c[i][1];
        out[blockIdx.x * blockDim.x + threadIdx.x] = s + x;
        return;
    }
    double x = 1.0 + lane * 1e-3;
    long long t0 = clock64();
    if (probe_on) {
        for (int it = 0; it < iters; it++) {
            double d = __shfl_sync(0xffffffffu, x, it & 31);
            double r; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
            double e = fma(-d, r, 1.0);
            double s = fma(e, e, e);
            x = fma(-r, s, x);
        }
    } else {
        while (clock64() - t0 < 2000000) {}
    }
    long long t1 = clock64();
    if (wid == 0 && lane == 0) { g_res[0] = t1 - t0; }
    __syncwarp();
    if (wid == 0 && lane == 0) stop = 1;
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}
template <int YIELD, int GROUPS>
void run(const char* name, double* out) {
    for (int probe = 0; probe <= 1; probe++) {
        const int iters = 2000;
        bench<YIELD, GROUPS><<<1, 32 * 12>>>(out, iters, probe);
        cudaDeviceSynchronize();
        long long r[4]; cudaMemcpyFromSymbol(r, g_res, 32);
        printf("%-44s probe=%d : %9.1f cycles per pivot step; stream: %.2f cycles per DMMA per warp (%s)\n", name, probe,
               probe ? (double)r[0] / iters : 0.0, (double)r[2] / (double)r[1], cudaGetErrorString(cudaGetLastError()));
    }
}
int main() {
    double* out; cudaMalloc(&out, 8 * 1024 * 4);
    run<0, 4>("no yield", out);
    run<1, 4>("nanosleep 0 per 16 DMMAs", out);
    run<5, 4>("nanosleep 20 per 16 DMMAs", out);
    run<2, 4>("3 dependent integer ops per 16 DMMAs", out);
    run<3, 4>("syncwarp per 16 DMMAs", out);
    run<4, 4>("dependent LDS feeding the next group", out);
    return 0;
}
