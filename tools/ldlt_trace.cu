// Developer tool: per-phase cycle breakdown of the batched LDL' factorisation (gf_ldlt.cu built with GF_LDLT_TRACE).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DGF_LDLT_TRACE -o tools/ldlt_trace tools/ldlt_trace.cu
//   tools/ldlt_trace [B] [N]
#include "../pygradflow_b200/csrc/gf_ldlt.cu"
#include <cstdio>
#include <cstdlib>
#include <vector>

__global__ void fill_kernel(double* K, int ld, int N, int npos) {
    double* Kb = K + (size_t)blockIdx.x * ld * ld;
    unsigned s0 = 1234567u + 7919u * blockIdx.x;
    for (int e = threadIdx.x; e < ld * ld; e += blockDim.x) {
        const int r = e / ld, c = e % ld;
        if (c > r) continue;
        unsigned h = s0 ^ (unsigned)(r * 73856093u) ^ (unsigned)(c * 19349663u);
        h ^= h >> 13; h *= 0x5bd1e995u; h ^= h >> 15;
        double v = ((h & 0xffffff) / (double)0x1000000) * 2.0 - 1.0;
        if (r == c) v = (r < npos ? 1.0 : -1.0) * (r < N ? N : 1.0);
        else if (r >= N) v = 0.0;
        Kb[(size_t)r * ld + c] = v;
    }
}

int main(int argc, char** argv) {
    const int B = argc > 1 ? atoi(argv[1]) : 4096, N = argc > 2 ? atoi(argv[2]) : 768;
    const int ld = ((N + 63) / 64) * 64, npos = N - N / 3;
    double *K, *dvec; int32_t *info, *nneg;
    cudaMalloc(&K, (size_t)B * ld * ld * 8); cudaMalloc(&dvec, (size_t)B * ld * 8);
    cudaMalloc(&info, B * 4); cudaMalloc(&nneg, B * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        fill_kernel<<<B, 256>>>(K, ld, N, npos);
        unsigned long long zero[64 * 2 * 16] = {0};
        cudaMemcpyToSymbol(g_ldlt_trace, zero, sizeof(zero));
        cudaMemcpyToSymbol(g_ldlt_fine, zero, 8 * sizeof(unsigned long long));
        cudaEventRecord(e0);
        int rc = gf_ldlt_factor(B, ld, N, nullptr, K, dvec, info, nneg, nullptr, nullptr, nullptr, B, nullptr);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
        if (rc) { printf("rc=%d\n", rc); return 1; }
    }
    std::vector<int32_t> h(B); cudaMemcpy(h.data(), info, B * 4, cudaMemcpyDeviceToHost);
    int bad = 0; for (int i = 0; i < B; i++) bad += h[i] != 0;
    printf("B=%d N=%d factor %.3f ms  %.2f TFLOP/s  bad=%d  err=%s\n", B, N, best, B * (double)N * N * N / 3 / best * 1e-9, bad,
           cudaGetErrorString(cudaGetLastError()));
    unsigned long long fine[8];
    cudaMemcpyFromSymbol(fine, g_ldlt_fine, sizeof(fine));
    printf("factor detail, thread 0, summed over all diagonal blocks and CTAs [Mcycles]: pivot block %.1f  barrier %.1f  panel %.1f  barrier %.1f  update %.1f\n",
           fine[0] * 1e-6, fine[1] * 1e-6, fine[2] * 1e-6, fine[3] * 1e-6, fine[4] * 1e-6);
    static unsigned long long t[64 * 2 * 16];
    cudaMemcpyFromSymbol(t, g_ldlt_trace, sizeof(t));
    printf("mean cycles per CTA.  chain(k): prologue mainloop epi-load trsm S-update | factor inverse writeback ; panel(k): prologue mainloop epi-load trsm\n");
    for (int k = 0; k < (N + 63) / 64; k++) {
        const unsigned long long* c = t + (k * 2 + 0) * 16; const unsigned long long* p = t + (k * 2 + 1) * 16;
        double nc = c[15] ? (double)c[15] : 1, np = p[15] ? (double)p[15] : 1;
        printf("k=%2d chain n=%6llu: %7.0f %7.0f %7.0f %7.0f %7.0f | %7.0f %7.0f %7.0f ; panel n=%6llu: %7.0f %7.0f %7.0f %7.0f\n", k, c[15],
               c[0] / nc, c[1] / nc, c[2] / nc, c[3] / nc, c[4] / nc, c[8] / nc, c[9] / nc, c[10] / nc, p[15], p[0] / np, p[1] / np,
               p[2] / np, p[3] / np);
    }
    return 0;
}
