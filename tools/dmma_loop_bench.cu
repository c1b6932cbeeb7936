// Microbenchmark: the LDL' panel main loop (32x32 warp tile, fragments from shared memory) with / without the
// in-loop DMUL that folds -D into the A fragment, at 1 and 2 CTAs (of 8 warps) per SM.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
constexpr int PKC = 32, PSP = 36;
template <bool DMUL, bool PIPE>
__global__ void __launch_bounds__(256, 2) loop_kernel(double* out, int chunks, int pad_smem) {
    extern __shared__ double sm[];
    double* As = sm; double* Bs = sm + 128 * PSP; double* Ds = Bs + 64 * PSP;
    for (int i = threadIdx.x; i < (128 + 64) * PSP + PKC; i += 256) sm[i] = 1.0 + 1e-6 * (i % 97);
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, wm = wid >> 1, wn = wid & 1, g = lane >> 2, q = lane & 3;
    double acc[4][4][2];
#pragma unroll
    for (int mi = 0; mi < 4; mi++)
#pragma unroll
        for (int ni = 0; ni < 4; ni++) { acc[mi][ni][0] = mi; acc[mi][ni][1] = ni; }
    const double* as = As + (wm * 32 + g) * PSP + q;
    const double* bs = Bs + (wn * 32 + g) * PSP + q;
    const double* ds = Ds + q;
    for (int c = 0; c < chunks; c++) {
        double a[4], bf[4], dcur;
#pragma unroll
        for (int mi = 0; mi < 4; mi++) a[mi] = as[mi * 8 * PSP];
#pragma unroll
        for (int ni = 0; ni < 4; ni++) bf[ni] = bs[ni * 8 * PSP];
        dcur = ds[0];
#pragma unroll
        for (int kk = 0; kk < PKC; kk += 4) {
            double an[4], bn[4], dn = 0.0;
            if (PIPE && kk + 4 < PKC) {
#pragma unroll
                for (int mi = 0; mi < 4; mi++) an[mi] = as[mi * 8 * PSP + kk + 4];
#pragma unroll
                for (int ni = 0; ni < 4; ni++) bn[ni] = bs[ni * 8 * PSP + kk + 4];
                dn = ds[kk + 4];
            }
            if (DMUL) {
#pragma unroll
                for (int mi = 0; mi < 4; mi++) a[mi] *= -dcur;
            }
#pragma unroll
            for (int mi = 0; mi < 4; mi++)
#pragma unroll
                for (int ni = 0; ni < 4; ni++) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], bf[ni]);
            if (kk + 4 < PKC) {
                if (PIPE) {
#pragma unroll
                    for (int mi = 0; mi < 4; mi++) a[mi] = an[mi];
#pragma unroll
                    for (int ni = 0; ni < 4; ni++) bf[ni] = bn[ni];
                    dcur = dn;
                } else {
#pragma unroll
                    for (int mi = 0; mi < 4; mi++) a[mi] = as[mi * 8 * PSP + kk + 4];
#pragma unroll
                    for (int ni = 0; ni < 4; ni++) bf[ni] = bs[ni * 8 * PSP + kk + 4];
                    dcur = ds[kk + 4];
                }
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int mi = 0; mi < 4; mi++)
#pragma unroll
        for (int ni = 0; ni < 4; ni++) s += acc[mi][ni][0] + acc[mi][ni][1];
    out[blockIdx.x * 256 + threadIdx.x] = s;
}
template <bool DMUL, bool PIPE>
void run(const char* name, int sms, double* out) {
    const int chunks = 4000;
    for (int per_sm = 1; per_sm <= 2; per_sm++) {
        int smem = per_sm == 1 ? 120 * 1024 : 100 * 1024;
        cudaFuncSetAttribute(loop_kernel<DMUL, PIPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        loop_kernel<DMUL, PIPE><<<sms * per_sm, 256, smem>>>(out, chunks, 0);
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        loop_kernel<DMUL, PIPE><<<sms * per_sm, 256, smem>>>(out, chunks, 0);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double fl = 512.0 * 16 * 8 * chunks * 8.0 * sms * per_sm;
        printf("%-28s %d CTA/SM: %7.2f TFLOP/s   (%s)\n", name, per_sm, fl / ms * 1e-9, cudaGetErrorString(cudaGetLastError()));
    }
}
int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double* out; cudaMalloc(&out, 8 * 256 * sms * 2);
    run<true, true>("DMUL + sw-pipelined LDS", sms, out);
    run<false, true>("no DMUL, sw-pipelined LDS", sms, out);
    run<true, false>("DMUL, plain LDS", sms, out);
    run<false, false>("no DMUL, plain LDS", sms, out);
    return 0;
}
