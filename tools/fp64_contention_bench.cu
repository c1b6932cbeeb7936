// Microbenchmark: latency of dependent FP64 / shuffle / DMMA steps in one warp per SM sub-partition while 0, 1 or 2
// other warps per sub-partition stream independent DMMAs (the situation of the LDL' diagonal-block factorisation
// next to a panel CTA).
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ long long g_res[8];
__device__ volatile int g_stop;
// warps [0, 4): probe warps (one per SMSP); warps [4, 4 + 4*nd): DMMA streamers
// LAST: the probe warps are the four HIGHEST warp indices of the block instead of the four lowest
template <int MODE, bool LAST = false>
__global__ void bench(double* out, int iters) {
    const int nw_ = blockDim.x >> 5;
    const int wid = LAST ? (int)(nw_ - 1 - (threadIdx.x >> 5)) : (int)(threadIdx.x >> 5), lane = threadIdx.x & 31;
    __shared__ volatile int stop;
    if (threadIdx.x == 0) stop = 0;
    __syncthreads();
    if (wid >= 4) {
        double c[16][2];
        double a = 1.0 + lane * 1e-9, b = 0.5;
#pragma unroll
        for (int i = 0; i < 16; i++) { c[i][0] = i; c[i][1] = -i; }
        while (!stop) {
#pragma unroll
            for (int r = 0; r < 4; r++)
#pragma unroll
                for (int i = 0; i < 16; i++) dmma884(c[i][0], c[i][1], a, b);
        }
        double s = 0;
#pragma unroll
        for (int i = 0; i < 16; i++) s += c[i][0] + c[i][1];
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;
        return;
    }
    double x = 1.0 + lane * 1e-3, y = 0.999, z = 1e-7, w0 = 0, w1 = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) x = fma(x, y, z);                                   // dependent DFMA
        if (MODE == 1) x = __shfl_sync(0xffffffffu, x, (lane + 1) & 31);   // dependent shuffle
        if (MODE == 2) { dmma884(w0, w1, x, y); }                          // dependent DMMA (accumulator chain)
        if (MODE == 3) { double r; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); x = r; }  // MUFU chain
        if (MODE == 4) {  // pivot step: shfl -> MUFU -> 3 fma
            double d = __shfl_sync(0xffffffffu, x, it & 31);
            double r; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
            double e = fma(-d, r, 1.0);
            double s = fma(e, e, e);
            x = fma(-r, s, x);
        }
        if (MODE == 6) {  // pivot step on the tensor pipe only: shfl -> MUFU -> 5 dependent DMMAs (k = 0 slot used)
            const bool q0 = (lane & 3) == 0;
            double d = __shfl_sync(0xffffffffu, x, it & 31);
            double r; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
            double e0 = 1.0, e1 = 1.0;
            dmma884(e0, e1, q0 ? -d : 0.0, q0 ? r : 0.0);          // e = 1 - d r
            double s0 = e0, s1 = e0;
            dmma884(s0, s1, q0 ? e0 : 0.0, q0 ? e0 : 0.0);         // sc = e + e e
            double t0_ = r, t1_ = r;
            dmma884(t0_, t1_, q0 ? r : 0.0, q0 ? s0 : 0.0);        // rj = r + r sc
            double m0 = 0.0, m1 = 0.0;
            dmma884(m0, m1, q0 ? y : 0.0, q0 ? t0_ : 0.0);         // w rj
            dmma884(w0, w1, q0 ? -y : 0.0, q0 ? m0 : 0.0);         // P -= w (w rj)
            x = w0 + 1.0;
        }
        if (MODE == 7) {  // two dependent DMMAs only
            dmma884(w0, w1, x, y);
            double m0 = 0.0, m1 = 0.0;
            dmma884(m0, m1, w0, y);
            x = m0;
        }
        if (MODE == 5) {  // 4 independent DFMAs then 1 dependent
            double a0 = fma(x, y, z), a1 = fma(x, z, y), a2 = fma(y, x, x), a3 = fma(z, x, x);
            x = (a0 + a1) + (a2 + a3);
        }
    }
    long long t1 = clock64();
    if (wid == 0 && lane == 0) { g_res[0] = t1 - t0; }
    __syncwarp();
    if (wid == 0 && lane == 0) stop = 1;
    out[blockIdx.x * blockDim.x + threadIdx.x] = x + w0 + w1;
}
template <int MODE, bool LAST = false>
void run(const char* name, double* out) {
    const int iters = 2000;
    for (int nd = 0; nd <= 3; nd++) {
        bench<MODE, LAST><<<1, 32 * (4 + 4 * nd)>>>(out, iters);
        cudaDeviceSynchronize();
        long long r; cudaMemcpyFromSymbol(&r, g_res, 8);
        printf("%-34s DMMA warps/SMSP=%d : %7.1f cycles per step  (%s)\n", name, nd, (double)r / iters, cudaGetErrorString(cudaGetLastError()));
    }
}
int main() {
    double* out; cudaMalloc(&out, 8 * 1024 * 4);
    run<0>("dependent DFMA", out);
    run<1>("dependent SHFL", out);
    run<2>("dependent DMMA", out);
    run<3>("dependent MUFU.RCP64H", out);
    run<4>("pivot step shfl+MUFU+3 DFMA", out);
    run<5>("4 indep DFMA + 3 DADD tree", out);
    run<6>("pivot step shfl+MUFU+5 dep DMMA", out);
    run<7>("2 dependent DMMA (through a operand)", out);
    run<0, true>("dependent DFMA, probe = LAST warps", out);
    run<2, true>("dependent DMMA, probe = LAST warps", out);
    run<4, true>("pivot step 3 DFMA, probe = LAST warps", out);
    return 0;
}
