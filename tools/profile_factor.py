"""Time the batched LDL' / LU factorisation + solve on synthetic quasi-definite KKT matrices (cfg5 style).

    python tools/profile_factor.py [--B 1024] [--N 768] [--method ldlt|lu] [--reps 3]
"""
import argparse, os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pygradflow_b200 import kernels as K

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=1024)
ap.add_argument("--N", type=int, default=768)
ap.add_argument("--method", default="ldlt")
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--check", action="store_true")
ap.add_argument("--general", action="store_true", help="dense random unsymmetric matrices (an interchange in almost every column)")
args = ap.parse_args()
B, N = args.B, args.N
dev = "cuda"
f64 = dict(dtype=torch.float64, device=dev)
m = N // 3
nI = N - m
ld = ((N + 63) // 64) * 64 if args.method == "ldlt" else N
gen = torch.Generator(device=dev); gen.manual_seed(0)
K0 = torch.zeros((B, ld, ld), **f64)
for lo in range(0, B, 128):
    hi = min(B, lo + 128)
    M = torch.randn((hi - lo, nI, nI), generator=gen, **f64)
    G = torch.bmm(M, M.transpose(1, 2)) / nI
    K0[lo:hi, :nI, :nI] = 0.5 * (G + G.transpose(1, 2)) + 1.1 * torch.eye(nI, **f64)
    A = torch.randn((hi - lo, m, nI), generator=gen, **f64)
    K0[lo:hi, nI:N, :nI] = A
    K0[lo:hi, :nI, nI:N] = A.transpose(1, 2)
    K0[lo:hi, nI:N, nI:N] = -0.99 * torch.eye(m, **f64)
    if ld > N:
        K0[lo:hi, N:, N:] = torch.eye(ld - N, **f64)
if args.general:
    K0[:, :N, :N] = torch.randn((B, N, N), generator=gen, **f64)
rhs0 = torch.randn((B, ld), generator=gen, **f64); rhs0[:, N:] = 0
Kw = torch.empty_like(K0); rhs = torch.empty_like(rhs0)
Nvec = torch.full((B,), N, dtype=torch.int32, device=dev)
info = torch.zeros((B,), dtype=torch.int32, device=dev)
w = K.WorkList.all(B)
if args.method == "ldlt":
    dvec = torch.zeros((B, ld), **f64); nneg = torch.zeros((B,), dtype=torch.int32, device=dev)
    npos = torch.full((B,), nI, dtype=torch.int32, device=dev)
    factor = lambda: K.ldlt_factor(Kw, N, Nvec, dvec, info, nneg, npos, w)
    solve = lambda: K.ldlt_solve(Kw, N, Nvec, rhs, w)
    flops = B * N ** 3 / 3.0
else:
    piv = torch.zeros((B, ld), dtype=torch.int32, device=dev)
    factor = lambda: K.lu_factor(Kw, N, Nvec, piv, info, w)
    solve = lambda: K.lu_solve(Kw, N, Nvec, piv, rhs, False, w)
    flops = B * 2 * N ** 3 / 3.0
tf, ts = [], []
for r in range(args.reps + 1):
    Kw.copy_(K0); rhs.copy_(rhs0)
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record(); factor(); e[1].record(); solve(); e[2].record()
    torch.cuda.synchronize()
    if r > 0:
        tf.append(e[0].elapsed_time(e[1])); ts.append(e[1].elapsed_time(e[2]))
fm, sm = min(tf), min(ts)
res = dict(method=args.method, B=B, N=N, factor_ms=fm, solve_ms=sm, factor_tflops=flops / fm * 1e-9,
           solve_gbs=B * N * N * 8 / sm * 1e-6, bad=int((info != 0).sum().item()))
if args.check:
    x = rhs[:, :N]
    Kd = K0[:, :N, :N]
    r = torch.bmm(Kd, x.unsqueeze(2)).squeeze(2) - rhs0[:, :N]
    res["max_residual"] = float(r.abs().max().item())
print(json.dumps(res))
