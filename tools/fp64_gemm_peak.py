"""Measure the FP64 GEMM peak (cuBLAS via torch.matmul) and HBM copy on this box -> gpurun_out/fp64_peak.json."""
import json, os, torch
dev = "cuda:0"
res = {}
for N in (4096, 8192):
    a = torch.randn(N, N, device=dev, dtype=torch.float64); b = torch.randn(N, N, device=dev, dtype=torch.float64)
    for _ in range(2): torch.matmul(a, b)
    best = 1e9
    for _ in range(5):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    res[f"dgemm_{N}_tflops"] = 2 * N**3 / best * 1e-9
    print(N, res[f"dgemm_{N}_tflops"], flush=True)
# batched dgemm 768
a = torch.randn(512, 768, 768, device=dev, dtype=torch.float64); b = torch.randn(512, 768, 768, device=dev, dtype=torch.float64)
for _ in range(2): torch.bmm(a, b)
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(); torch.bmm(a, b); e1.record(); torch.cuda.synchronize()
res["bmm_512x768_tflops"] = 2 * 512 * 768**3 / e0.elapsed_time(e1) * 1e-9
# batched LU through torch (cuSOLVER/MAGMA) as a library reference point
K = torch.randn(256, 768, 768, device=dev, dtype=torch.float64); K = K @ K.transpose(1, 2) + 768 * torch.eye(768, device=dev, dtype=torch.float64)
torch.linalg.lu_factor(K)
e0.record(); torch.linalg.lu_factor(K); e1.record(); torch.cuda.synchronize()
res["torch_lu_factor_256x768_ms"] = e0.elapsed_time(e1)
torch.linalg.cholesky(K)
e0.record(); torch.linalg.cholesky(K); e1.record(); torch.cuda.synchronize()
res["torch_cholesky_256x768_ms"] = e0.elapsed_time(e1)
print(res)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/fp64_peak.json", "w"), indent=1)
