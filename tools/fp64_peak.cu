// FP64 pipe microbenchmark for B200 (sm_100a): DFMA vs DMMA (mma.sync m8n8k4 / m16n8k8 f64) throughput.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dfma_kernel(double* out, int iters) {
    double a[16];
    double x = 1.0 + threadIdx.x * 1e-9, y = 0.999999;
#pragma unroll
    for (int i = 0; i < 16; i++) a[i] = i * 0.5;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) a[i] = fma(a[i], x, y);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__global__ void dmma_kernel(double* out, int iters) {
    double c[16][2];
    double a = 1.0 + threadIdx.x * 1e-9, b = 0.5;
#pragma unroll
    for (int i = 0; i < 16; i++) { c[i][0] = i; c[i][1] = -i; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) dmma884(c[i][0], c[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void dmma1688(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}

__global__ void dmma1688_kernel(double* out, int iters) {
    double c[8][4];
    double a[4] = {1.0 + threadIdx.x * 1e-9, 0.5, 0.25, 0.125}, b[2] = {0.5, 0.75};
#pragma unroll
    for (int i = 0; i < 8; i++) { c[i][0] = i; c[i][1] = -i; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) dmma1688(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_it(F f) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f();
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    f();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int threads = 256, blocks = sms * 8, iters = 20000;
    double* out; cudaMalloc(&out, sizeof(double) * threads * blocks);
    for (int warps = 1; warps <= 8; warps *= 2) {
        int th = 32 * warps * 4;  // warps per SMSP * 4 SMSPs (1 block per SM)
        if (th > 1024) break;
        float ms = time_it([&] { dfma_kernel<<<sms, th>>>(out, iters); });
        double fl = 2.0 * 16 * iters * (double)th * sms;
        printf("DFMA       1 CTA/SM x %4d thr: %8.2f TFLOP/s\n", th, fl / ms * 1e-9);
        ms = time_it([&] { dmma_kernel<<<sms, th>>>(out, iters); });
        fl = 512.0 * 16 * iters * (double)(th / 32) * sms;
        printf("DMMA m8n8k4  1 CTA/SM x %4d thr: %8.2f TFLOP/s\n", th, fl / ms * 1e-9);
        ms = time_it([&] { dmma1688_kernel<<<sms, th>>>(out, iters); });
        fl = 2048.0 * 8 * iters * (double)(th / 32) * sms;
        printf("DMMA m16n8k8 1 CTA/SM x %4d thr: %8.2f TFLOP/s\n", th, fl / ms * 1e-9);
    }
    float ms = time_it([&] { dfma_kernel<<<blocks, threads>>>(out, iters); });
    printf("DFMA full occupancy: %8.2f TFLOP/s\n", 2.0 * 16 * iters * (double)threads * blocks / ms * 1e-9);
    ms = time_it([&] { dmma_kernel<<<blocks, threads>>>(out, iters); });
    printf("DMMA m8n8k4 full occupancy: %8.2f TFLOP/s\n", 512.0 * 16 * iters * (double)(threads / 32) * blocks / ms * 1e-9);
    ms = time_it([&] { dmma1688_kernel<<<blocks, threads>>>(out, iters); });
    printf("DMMA m16n8k8 full occupancy: %8.2f TFLOP/s\n", 2048.0 * 8 * iters * (double)(threads / 32) * blocks / ms * 1e-9);
    return 0;
}
