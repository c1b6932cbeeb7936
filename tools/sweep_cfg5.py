"""cfg5 of BASELINE.json: KKT factor + solve sweep N = 32..2048 x batch 64..16384 against the reference's CPU
linear solver (scipy splu = SuperLU exactly as pygradflow/linear_solver/lu_solver.py:14,21).

    python tools/sweep_cfg5.py [--out profiles/r01_cfg5_sweep.json] [--cpu-samples 4] [--methods ldlt,lu]

Matrices: quasi-definite K = [[H + lamb I, A'], [A, -delta I]] of order N with m = N // 3 (synth.kkt_instance's
family, generated on the device).  Timing: CUDA events, best of `--reps`, inputs (>= 134 MB except the smallest
cases, which are repeated over distinct copies) rewritten before every repetition, so nothing is L2-resident.
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from pygradflow_b200 import kernels as K

NS = [32, 64, 128, 256, 512, 1024, 2048]
BS = [64, 256, 1024, 4096, 16384]
MAX_BYTES = 24e9


def make_batch(B, N, ld, gen, dev):
    f64 = dict(dtype=torch.float64, device=dev)
    m = N // 3
    nI = N - m
    K0 = torch.zeros((B, ld, ld), **f64)
    step = max(1, min(B, int(2e8 // max(1, nI * nI))))
    for lo in range(0, B, step):
        hi = min(B, lo + step)
        M = torch.randn((hi - lo, nI, nI), generator=gen, **f64)
        G = torch.bmm(M, M.transpose(1, 2)) / max(nI, 1)
        K0[lo:hi, :nI, :nI] = 0.5 * (G + G.transpose(1, 2)) + 1.1 * torch.eye(nI, **f64)
        if m > 0:
            A = torch.randn((hi - lo, m, nI), generator=gen, **f64)
            K0[lo:hi, nI:N, :nI] = A
            K0[lo:hi, :nI, nI:N] = A.transpose(1, 2)
            K0[lo:hi, nI:N, nI:N] = -0.99 * torch.eye(m, **f64)
        if ld > N:
            K0[lo:hi, N:, N:] = torch.eye(ld - N, **f64)
    rhs0 = torch.randn((B, ld), generator=gen, **f64)
    rhs0[:, N:] = 0
    return K0, rhs0, nI


def run_case(method, B, N, reps, dev):
    ld = max(((N + 63) // 64) * 64, 64) if method == "ldlt" else N
    gen = torch.Generator(device=dev)
    gen.manual_seed(4000 + N)
    K0, rhs0, nI = make_batch(B, N, ld, gen, dev)
    Kw, rhs = torch.empty_like(K0), torch.empty_like(rhs0)
    i32 = dict(dtype=torch.int32, device=dev)
    Nvec = torch.full((B,), N, **i32)
    info = torch.zeros((B,), **i32)
    w = K.WorkList.all(B)
    if method == "ldlt":
        dvec = torch.zeros((B, ld), dtype=torch.float64, device=dev)
        nneg = torch.zeros((B,), **i32)
        npos = torch.full((B,), nI, **i32)
        factor = lambda: K.ldlt_factor(Kw, N, Nvec, dvec, info, nneg, npos, w)
        solve = lambda: K.ldlt_solve(Kw, N, Nvec, rhs, w)
        flops = B * N ** 3 / 3.0
    else:
        piv = torch.zeros((B, ld), **i32)
        factor = lambda: K.lu_factor(Kw, N, Nvec, piv, info, w)
        solve = lambda: K.lu_solve(Kw, N, Nvec, piv, rhs, False, w)
        flops = B * 2 * N ** 3 / 3.0
    tf, ts = [], []
    for r in range(reps + 1):
        Kw.copy_(K0)
        rhs.copy_(rhs0)
        torch.cuda.synchronize()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        factor()
        e[1].record()
        solve()
        e[2].record()
        torch.cuda.synchronize()
        if r > 0:
            tf.append(e[0].elapsed_time(e[1]))
            ts.append(e[1].elapsed_time(e[2]))
    S = min(B, 8)
    x = rhs[:S, :N]
    res = torch.bmm(K0[:S, :N, :N].tril() + K0[:S, :N, :N].tril(-1).transpose(1, 2), x.unsqueeze(2)).squeeze(2) - rhs0[:S, :N]
    fm, sm = min(tf), min(ts)
    out = dict(method=method, B=B, N=N, factor_ms=fm, solve_ms=sm, factor_gflops=flops / fm * 1e-6,
               solve_gbs=B * (N * N + 2 * N) * 8 / sm * 1e-6, factor_solve_per_s=B / ((fm + sm) * 1e-3),
               bad=int((info != 0).sum().item()), max_residual=float(res.abs().max().item()))
    host = K0[:4, :N, :N].cpu().numpy(), rhs0[:4, :N].cpu().numpy(), x[:4].cpu().numpy()
    del K0, Kw, rhs0, rhs
    torch.cuda.empty_cache()
    return out, host


def cpu_case(mats, rhss, sols):
    """LUSolver of the reference: splu(csc_matrix(K)) + solve, one core."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl

    tf, ts, err = [], [], 0.0
    for Km, r, xg in zip(mats, rhss, sols):
        Km = np.tril(Km) + np.tril(Km, -1).T
        A = sp.csc_matrix(Km)
        t0 = time.perf_counter()
        lu = spl.splu(A)
        t1 = time.perf_counter()
        x = lu.solve(r)
        t2 = time.perf_counter()
        tf.append(t1 - t0)
        ts.append(t2 - t1)
        err = max(err, float(np.max(np.abs(x - xg)) / max(1.0, np.max(np.abs(x)))))
    return dict(factor_ms=1e3 * min(tf), solve_ms=1e3 * min(ts), gpu_vs_splu_rel=err)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="profiles/r01_cfg5_sweep.json")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--methods", default="ldlt,lu")
    ap.add_argument("--cpu-samples", type=int, default=3)
    ap.add_argument("--ns", default=",".join(map(str, NS)))
    ap.add_argument("--bs", default=",".join(map(str, BS)))
    args = ap.parse_args()
    dev = "cuda"
    rows, cpu = [], {}
    for N in [int(v) for v in args.ns.split(",")]:
        for B in [int(v) for v in args.bs.split(",")]:
            ldmax = ((N + 63) // 64) * 64
            if 2.0 * B * ldmax * ldmax * 8 > MAX_BYTES:
                continue
            for method in args.methods.split(","):
                if method == "lu" and N >= 1024 and B > 256:
                    continue  # the pivoted fallback path: sampled, not swept, at the large orders
                row, host = run_case(method, B, N, args.reps, dev)
                if N not in cpu and args.cpu_samples > 0:
                    k = args.cpu_samples
                    cpu[N] = cpu_case(host[0][:k], host[1][:k], host[2][:k])
                if N in cpu:
                    c = cpu[N]
                    row["cpu_splu_factor_ms"] = c["factor_ms"]
                    row["cpu_splu_solve_ms"] = c["solve_ms"]
                    row["speedup_vs_one_core"] = (c["factor_ms"] + c["solve_ms"]) * B / (row["factor_ms"] + row["solve_ms"])
                    if method == "ldlt":
                        row["gpu_vs_splu_rel"] = c["gpu_vs_splu_rel"]
                rows.append(row)
                print(json.dumps(row), flush=True)
    with open(args.out, "w") as f:
        json.dump(dict(rows=rows, cpu=cpu, device=torch.cuda.get_device_name(0)), f, indent=1)
    # markdown table
    md = ["| N | B | method | factor ms | TFLOP/s | solve ms | solve GB/s | factor+solve /s | x one CPU core (splu) |", "|---|---|---|---|---|---|---|---|---|"]
    for r in rows:
        md.append(f"| {r['N']} | {r['B']} | {r['method']} | {r['factor_ms']:.3f} | {r['factor_gflops'] * 1e-3:.2f} | "
                  f"{r['solve_ms']:.3f} | {r['solve_gbs']:.0f} | {r['factor_solve_per_s']:.0f} | "
                  f"{r.get('speedup_vs_one_core', float('nan')):.0f} |")
    with open(os.path.splitext(args.out)[0] + ".md", "w") as f:
        f.write("\n".join(md) + "\n")


if __name__ == "__main__":
    main()
