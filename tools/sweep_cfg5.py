"""cfg5 of BASELINE.json: KKT factor + solve sweep N = 32..2048 x batch 64..16384 against the reference's CPU
linear solver (scipy splu = SuperLU exactly as pygradflow/linear_solver/lu_solver.py:14,21).

    python tools/sweep_cfg5.py [--out profiles/r02_cfg5_sweep_1.json] [--cpu-samples 4] [--methods ldlt,lu]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port 29533 \
        tools/sweep_cfg5.py --out profiles/r02_cfg5_sweep_G.json

Under torchrun the batch of every case is block-sharded over the G ranks (instances are independent: no collective on
the data path; rank r factorises and solves B / G matrices), all ranks start a case together (barrier) and the case's
time is the MAXIMUM over the ranks (one all_reduce of the two CUDA-event times) -- factor+solve/s is the whole job's.
Cases: B x N^2 x 8 bytes <= 64 GB per GPU; the tool keeps the pristine copy next to the factorised one.

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
MAX_BYTES = 64e9  # B N^2 8 bytes <= 64 GB per GPU


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


def run_case(method, B, N, reps, dev, rank=0, world=1):
    """B = this rank's share of the case's batch."""
    ld = max(((N + 63) // 64) * 64, 64) if method == "ldlt" else N
    gen = torch.Generator(device=dev)
    gen.manual_seed(4000 + N + 7919 * rank)
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
        if world > 1:
            torch.distributed.barrier()
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
    bad = int((info != 0).sum().item())
    maxres = float(res.abs().max().item())
    if world > 1:  # per repetition the slowest rank counts; then the best repetition
        t = torch.tensor([tf, ts], dtype=torch.float64, device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        tf, ts = t[0].tolist(), t[1].tolist()
        agg = torch.tensor([float(bad), maxres], dtype=torch.float64, device=dev)
        torch.distributed.all_reduce(agg, op=torch.distributed.ReduceOp.MAX)
        bad, maxres = int(agg[0].item()), float(agg[1].item())
    fm, sm = min(tf), min(ts)
    Bt = B * world
    out = dict(method=method, B=Bt, N=N, gpus=world, factor_ms=fm, solve_ms=sm, factor_gflops=flops * world / fm * 1e-6,
               solve_gbs=Bt * (N * N + 2 * N) * 8 / sm * 1e-6, factor_solve_per_s=Bt / ((fm + sm) * 1e-3),
               bad=bad, max_residual=maxres)
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
    ap.add_argument("--out", default="profiles/r02_cfg5_sweep_1.json")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--methods", default="ldlt,lu")
    ap.add_argument("--cpu-samples", type=int, default=3)
    ap.add_argument("--ns", default=",".join(map(str, NS)))
    ap.add_argument("--bs", default=",".join(map(str, BS)))
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        torch.distributed.init_process_group("nccl")
    dev = "cuda"
    ref = {}
    ref_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles",
                            "r02_cfg5_reference_lusolver.json")
    if os.path.exists(ref_path):  # the REAL reference's LUSolver, timed in the build container (tools/cfg5_reference_cpu.py)
        ref = json.load(open(ref_path))["rows"]
    rows, cpu = [], {}
    for N in [int(v) for v in args.ns.split(",")]:
        for B in [int(v) for v in args.bs.split(",")]:
            ldmax = ((N + 63) // 64) * 64
            if 1.0 * (B // world) * ldmax * ldmax * 8 > MAX_BYTES or B % world != 0:
                continue
            for method in args.methods.split(","):
                if method == "lu" and N >= 1024 and B // world > 1024:
                    continue  # the pivoted path: sampled, not swept, at the large orders
                row, host = run_case(method, B // world, N, args.reps, dev, rank, world)
                if rank != 0:
                    continue
                if str(N) in ref:
                    r = ref[str(N)]
                    row["reference_lusolver_factor_ms"] = r["factor_ms"]
                    row["reference_lusolver_solve_ms"] = r["solve_ms"]
                    row["speedup_vs_reference_one_core"] = (r["factor_ms"] + r["solve_ms"]) * row["B"] / (
                        row["factor_ms"] + row["solve_ms"])
                if N not in cpu and args.cpu_samples > 0:
                    k = args.cpu_samples
                    cpu[N] = cpu_case(host[0][:k], host[1][:k], host[2][:k])
                if N in cpu:
                    c = cpu[N]
                    row["cpu_splu_factor_ms"] = c["factor_ms"]
                    row["cpu_splu_solve_ms"] = c["solve_ms"]
                    row["speedup_vs_one_core"] = (c["factor_ms"] + c["solve_ms"]) * row["B"] / (row["factor_ms"] + row["solve_ms"])
                    if method == "ldlt":
                        row["gpu_vs_splu_rel"] = c["gpu_vs_splu_rel"]
                rows.append(row)
                print(json.dumps(row), flush=True)
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    if rank != 0:
        return
    with open(args.out, "w") as f:
        json.dump(dict(rows=rows, cpu=cpu, gpus=world, device=torch.cuda.get_device_name(0),
                       host_cores=os.cpu_count()), f, indent=1)
    # markdown table
    md = [f"cfg5 sweep on {world} x {torch.cuda.get_device_name(0)}: batch block-sharded over the GPUs, time = max over ranks; "
          "TFLOP/s and GB/s are the whole job's (N^3/3 flop per LDL', 2N^3/3 per LU; 8(N^2+2N) bytes per solve).", "",
          "| N | B | GPUs | method | factor ms | TFLOP/s | solve ms | solve GB/s | factor+solve /s | x one core, splu on the box | x one core, reference LUSolver (build container) |",
          "|---|---|---|---|---|---|---|---|---|---|---|"]
    for r in rows:
        md.append(f"| {r['N']} | {r['B']} | {r['gpus']} | {r['method']} | {r['factor_ms']:.3f} | {r['factor_gflops'] * 1e-3:.2f} | "
                  f"{r['solve_ms']:.3f} | {r['solve_gbs']:.0f} | {r['factor_solve_per_s']:.0f} | "
                  f"{r.get('speedup_vs_one_core', float('nan')):.0f} | {r.get('speedup_vs_reference_one_core', float('nan')):.0f} |")
    with open(os.path.splitext(args.out)[0] + ".md", "w") as f:
        f.write("\n".join(md) + "\n")


if __name__ == "__main__":
    main()
