// Developer tool: per-phase cycle breakdown of the stage-structured KKT kernels (gf_blocktri.cu built with GF_STAGE_TRACE).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DGF_STAGE_TRACE -o tools/stage_trace_bin tools/stage_trace.cu
//   tools/stage_trace_bin [B] [S]
#include "../pygradflow_b200/csrc/gf_blocktri.cu"
#include <cstdio>
#include <cstdlib>
#include <vector>

__global__ void fill(double* Jc, double* Hd, double* F, uint8_t* act, double* dt, double* rho, int S, int nu) {
    const int w = 8 + nu, n = S * w, m = S * 8, jw = 8 + w, b = blockIdx.x;
    unsigned s0 = 1234567u + 7919u * b;
    auto rnd = [&](unsigned i) { unsigned h = s0 ^ (i * 73856093u); h ^= h >> 13; h *= 0x5bd1e995u; h ^= h >> 15; return ((h & 0xffffff) / (double)0x1000000) * 2.0 - 1.0; };
    for (int e = threadIdx.x; e < m * jw; e += blockDim.x) {
        const int j = e / (8 * jw), c = (e % (8 * jw)) / 8, r = e % 8;  // Jc[j][c][r]
        double v = 0.3 * rnd(e);
        if (c < 8) v = j == 0 ? 0.0 : -( (r == c ? 1.0 : 0.0) + 0.05 * v);
        else if (c < 16) v = (r == c - 8) ? 1.0 : 0.0;
        Jc[(size_t)b * m * jw + e] = v;
    }
    for (int i = threadIdx.x; i < n; i += blockDim.x) { Hd[(size_t)b * n + i] = 1.0 + 0.5 * rnd(1000000 + i); act[(size_t)b * n + i] = (i % w) >= 8 && (i % 5 == 0); }
    for (int i = threadIdx.x; i < n + m; i += blockDim.x) F[(size_t)b * (n + m) + i] = rnd(2000000 + i);
    if (threadIdx.x == 0) { dt[b] = 0.5; rho[b] = 1e-3; }
}

int main(int argc, char** argv) {
    const int B = argc > 1 ? atoi(argv[1]) : 128, S = argc > 2 ? atoi(argv[2]) : 128, nu = 8;
    const int w = 8 + nu, n = S * w, m = S * 8, jw = 8 + w;
    double *Jc, *Hd, *F, *dt, *rho, *Ti, *Lc, *Uc, *sol; uint8_t* act; int32_t *info, *nneg;
    cudaMalloc(&Jc, (size_t)B * m * jw * 8); cudaMalloc(&Hd, (size_t)B * n * 8); cudaMalloc(&F, (size_t)B * (n + m) * 8);
    cudaMalloc(&act, (size_t)B * n); cudaMalloc(&dt, B * 8); cudaMalloc(&rho, B * 8);
    cudaMalloc(&Ti, (size_t)B * S * 64 * 8); cudaMalloc(&Lc, (size_t)B * S * 64 * 8); cudaMalloc(&Uc, (size_t)B * S * 64 * 8);
    cudaMalloc(&sol, (size_t)B * (n + m) * 8); cudaMalloc(&info, B * 4); cudaMalloc(&nneg, B * 4);
    fill<<<B, 256>>>(Jc, Hd, F, act, dt, rho, S, nu);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float bf = 1e30f, bs = 1e30f;
    unsigned long long zero[32] = {0};
    const int reps = 5;
    for (int rep = 0; rep < reps; rep++) {
        if (rep == reps - 1) cudaMemcpyToSymbol(g_stage_trace, zero, sizeof(zero));
        cudaEventRecord(e0);
        int rc = gf_stage_kkt_factor(B, S, 8, nu, Jc, Hd, act, dt, rho, Ti, Lc, Uc, info, nneg, nullptr, nullptr, B, nullptr);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < bf) bf = ms;
        cudaEventRecord(e0);
        rc |= gf_stage_kkt_solve(B, S, 8, nu, Jc, Hd, act, F, dt, rho, Ti, Lc, Uc, sol, n + m, nullptr, nullptr, B, nullptr);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1); if (ms < bs) bs = ms;
        if (rc) { printf("rc=%d\n", rc); return 1; }
    }
    std::vector<int32_t> h(B); cudaMemcpy(h.data(), info, B * 4, cudaMemcpyDeviceToHost);
    int bad = 0; for (int i = 0; i < B; i++) bad += h[i] != 0;
    unsigned long long t[32]; cudaMemcpyFromSymbol(t, g_stage_trace, sizeof(t));
    printf("B=%d S=%d factor %.1f us solve %.1f us bad=%d err=%s\n", B, S, bf * 1e3, bs * 1e3, bad, cudaGetErrorString(cudaGetLastError()));
    printf("factor cycles/CTA: invD %.0f  formM %.0f  CR-invert %.0f  CR-update %.0f  last %.0f\n", t[0] / (double)B, t[1] / (double)B, t[2] / (double)B, t[3] / (double)B, t[4] / (double)B);
    printf("solve cycles/CTA: rhs %.0f  forward %.0f  backward %.0f  output %.0f\n", t[8] / (double)B, t[9] / (double)B, t[10] / (double)B, t[11] / (double)B);
    return 0;
}
