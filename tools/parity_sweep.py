"""Full-size parity sweep: BatchedSolver (CUDA path) against the CPU oracle (scipy splu, the reference's arithmetic)
on identical inputs, instance by instance and outer iteration by outer iteration.

    python tools/parity_sweep.py [--cfg 2,3,4] [--count 128,128,8] [--out profiles/r02_parity_sweep.json]

For every instance the two sides are compared on: termination status, iteration count, accepted-step count, the
accept / reject / failure sequence, the active set of every outer iteration (64-bit signature), and the final
iterate; the lambda trajectory is reported (maximal relative deviation while the decisions agree), not gated.  The first outer iteration where anything differs is reported with
the contraction ratio theta there, the active-set margin |p - (bound -/+ 1e-8)| of the oracle at that iteration
("active-set bit-consistency") and whether the instance had already passed its rounding-noise horizon
(first iteration with theta < 1e-8: from there on the reference feeds log(theta) of pure rounding noise into its PI
controller, distance_ratio_control.py:57-63, and lambda depends on the last bits of the linear solve).

The oracle is test infrastructure; nothing under pygradflow_b200/ imports it.
"""
import argparse
import json
import os
import sys
import time
from concurrent.futures import ProcessPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

THETA_NOISE = 1e-8
# lambda_next = lambda / exp(K_P e + K_I sum e), e = log(theta_ref) - log(theta), theta = |d2| / |d1|: the relative error of
# lambda is 0.2 x that of theta, and |d2| -- the length of the second Newton step, 1e-6 .. 1e-12 of |d1| near convergence
# -- carries the absolute rounding error of the linear solve (~1e-16 |x|), i.e. up to 1e-4 relative.  The lambda
# trajectories of two correct implementations therefore agree to 1e-6 .. 1e-5 at best; decisions (accept / reject,
# active sets, termination, iteration counts) are what must be identical, lambda is reported only.
LAMB_RTOL = 1e-6
HASH_MUL = 2654435761


def _weights(n):
    return ((np.arange(1, n + 1, dtype=np.int64) * HASH_MUL) % (1 << 31)).astype(np.int64)


def _instance_data(cfg, k, stages):
    from pygradflow_b200 import synth

    if cfg == 2:
        return synth.rosenbrock_instance(k, 64)
    if cfg == 3:
        return synth.qp_instance(k, 512, 256)
    return synth.ocp_instance(k, stages=stages)


def _oracle_solve(args):
    """One instance through the oracle's Solver.solve; returns compact per-iteration records."""
    cfg, k, stages, limit = args
    from threadpoolctl import threadpool_limits

    threadpool_limits(1)  # one BLAS thread per worker process, like the reference's runner (one process per core)
    from oracle import gradflow_oracle as orc

    d = _instance_data(cfg, k, stages)
    prm = orc.OracleParams(iteration_limit=limit)
    if cfg == 2:
        p = orc.ChainedRosenbrock(d["a"], d["b"], d["lb"], d["ub"])
    elif cfg == 3:
        p = orc.DenseQP(d["H"], d["A"], d["g"], d["b"], d["lb"], d["ub"])
    else:
        p = orc.OCP(d["A"], d["B"], d["Q"], d["R"], d["xinit"], d["umax"], d["h"], sparse=True)
        prm.sparse = True
    margins = []

    def hook(method, kstep, result):
        if kstep == 0:  # margin of the active-set decision of this outer iteration (implicit_func.py:44,246)
            f = method.func
            pp = f.projection_initial(method.orig_iterate, method.rho, method.tau)
            with np.errstate(invalid="ignore"):
                mg = np.minimum(np.abs(pp - (f.lb - 1e-8)), np.abs(pp - (f.ub + 1e-8)))
            mg = mg[np.isfinite(mg)]
            margins.append(float(mg.min()) if mg.size else float("inf"))

    t0 = time.perf_counter()
    failed = None
    try:
        res = orc.Solver(p, prm).solve(d["x0"], d["y0"], record=True, step_hook=hook)
    except orc.LambMaxError as err:
        failed = str(err)
    cpu_s = time.perf_counter() - t0
    if failed is not None:
        return dict(k=k, error=failed, cpu_s=cpu_s)
    w = _weights(p.num_vars)
    tr = res.trace
    code = np.array([5 if t["newton_steps"] == 0 else (2 if (t["accept"] and t["newton_steps"] == 1) else
                                                        (3 if t["accept"] else 4)) for t in tr], dtype=np.int8)
    ah = np.array([-1 if t["active"] is None else int(np.dot(t["active"].astype(np.int64), w)) for t in tr],
                  dtype=np.int64)
    # margins: one per outer iteration whose first Newton step ran (a failed factorisation raises before the hook)
    mg = np.full(len(tr), np.nan)
    j = 0
    for i, t in enumerate(tr):
        if t["newton_steps"] > 0 and j < len(margins):
            mg[i] = margins[j]
            j += 1
    return dict(k=k, status=int(res.status), iterations=int(res.iterations), accepted=int(res.accepted_steps),
                code=code, lamb=np.array([t["lamb_next"] for t in tr]), theta=np.array([t["theta"] for t in tr]),
                ahash=ah, margin=mg, x=res.x, y=res.y, cpu_s=cpu_s)


def _gpu_solve(cfg, ks, stages, limit):
    import torch

    from pygradflow_b200 import synth
    from pygradflow_b200.params import Params
    from pygradflow_b200.problem import BatchedOCP, BatchedQP, BatchedRosenbrock
    from pygradflow_b200.solver import BatchedSolver

    if cfg == 2:
        d = synth.rosenbrock_batch(ks, 64)
        prob = BatchedRosenbrock(d["a"], d["b"], d["lb"], d["ub"])
        x0, y0 = d["x0"], None
    elif cfg == 3:
        d = synth.qp_batch(ks, 512, 256)
        prob = BatchedQP(d["H"], d["A"], d["g"], d["b"], d["lb"], d["ub"])
        x0, y0 = d["x0"], d["y0"]
    else:
        d = synth.ocp_batch(ks, stages=stages)
        prob = BatchedOCP(d["A"], d["B"], d["Q"], d["R"], d["xinit"], d["umax"], d["h"])
        x0, y0 = d["x0"], d["y0"]
    B = prob.B
    solver = BatchedSolver(prob, Params(iteration_limit=limit))
    w = torch.as_tensor(_weights(prob.n), device=prob.device)
    rec = [dict(code=[], lamb=[], theta=[], ahash=[]) for _ in range(B)]

    def hook(outer, s):
        st = s.status.cpu().numpy()
        ph = s.phase.cpu().numpy()
        ln = s.lamb_next.cpu().numpy()
        th = s.theta.cpu().numpy()
        ah = (s.engine.active.to(torch.int64) * w[None, :]).sum(dim=1).cpu().numpy()
        for b in range(B):
            if st[b] != 0 or ph[b] == 0:
                continue
            r = rec[b]
            r["code"].append(int(ph[b]))
            r["lamb"].append(float(ln[b]))
            r["theta"].append(float(th[b]) if ph[b] in (3, 4) else float("nan"))
            r["ahash"].append(-1 if ph[b] == 5 else int(ah[b]))

    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = solver.solve(x0, y0, on_iteration=hook)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    out = []
    for b in range(B):
        r = rec[b]
        out.append(dict(status=int(res.status[b].item()), iterations=int(res.iterations[b].item()),
                        accepted=int(res.accepted_steps[b].item()), code=np.array(r["code"], dtype=np.int8),
                        lamb=np.array(r["lamb"]), theta=np.array(r["theta"]), ahash=np.array(r["ahash"], dtype=np.int64),
                        x=res.x[b].cpu().numpy(), y=res.y[b].cpu().numpy()))
    return out, wall, solver.engine.linear.name


def _rel(a, b):
    if a.size == 0:
        return 0.0
    return float(np.max(np.abs(a - b)) / max(1.0, float(np.max(np.abs(b)))))


def compare_instance(g, c):
    """g: GPU record, c: oracle record -> dict of findings."""
    if "error" in c:
        return dict(k=c["k"], oracle_error=c["error"], status_gpu=g["status"], identical=g["status"] == 6,
                    pre_horizon=True, first_div=None)
    n = min(len(g["code"]), len(c["code"]))
    th = c["theta"]
    noisy = np.where((th == th) & (th < THETA_NOISE))[0]
    horizon = int(noisy[0]) if noisy.size else len(th)
    first, why = None, None
    for i in range(n):
        if g["code"][i] != c["code"][i]:
            first, why = i, "accept"
        elif g["ahash"][i] != c["ahash"][i]:
            first, why = i, "active_set"
        if first is not None:
            break
    if first is None and len(g["code"]) != len(c["code"]):
        first, why = n, "length"
    lam_rel, lam_sep = 0.0, None
    upto = n if first is None else first
    if upto > 0:
        lr = np.abs(g["lamb"][:upto] - c["lamb"][:upto]) / np.abs(c["lamb"][:upto])
        lam_rel = float(np.max(lr))
        sep = np.where(lr > LAMB_RTOL)[0]
        lam_sep = int(sep[0]) if sep.size else None   # informational: first iteration with |dlambda| / lambda > LAMB_RTOL
    out = dict(k=c["k"], status_gpu=g["status"], status_cpu=c["status"], iters_gpu=g["iterations"],
               iters_cpu=c["iterations"], accepted_gpu=g["accepted"], accepted_cpu=c["accepted"], horizon=horizon,
               hit_horizon=bool(noisy.size), first_div=first, div_kind=why, lamb_rel_before_div=lam_rel,
               lamb_separation_iter=lam_sep,
               x_rel=_rel(g["x"], c["x"]), y_rel=_rel(g["y"], c["y"]), cpu_s=c["cpu_s"])
    if first is not None and first < len(c["theta"]):
        out["theta_at_div"] = float(c["theta"][first]) if c["theta"][first] == c["theta"][first] else None
        out["margin_at_div"] = float(c["margin"][first]) if c["margin"][first] == c["margin"][first] else None
    out["identical"] = bool(first is None and g["status"] == c["status"] and g["iterations"] == c["iterations"]
                            and g["accepted"] == c["accepted"])
    out["pre_horizon"] = bool(first is None or first <= horizon)   # the difference (if any) shows up before / at the horizon
    return out


def sweep(cfg, count, stages=128, limit=None, workers=None):
    ks = list(range(count))
    gpu, wall, linear = _gpu_solve(cfg, ks, stages, limit)
    workers = workers or os.cpu_count() or 1
    t0 = time.perf_counter()
    with ProcessPoolExecutor(max_workers=min(workers, count)) as ex:
        cpu = list(ex.map(_oracle_solve, [(cfg, k, stages, limit) for k in ks]))
    cpu_wall = time.perf_counter() - t0
    rows = [compare_instance(g, c) for g, c in zip(gpu, cpu)]
    ident = sum(r["identical"] for r in rows)
    never = [r for r in rows if not r.get("hit_horizon", False)]
    post = [r for r in rows if r.get("hit_horizon", False)]
    strict_fail = [r for r in rows if not r["identical"] and r["first_div"] is not None and r["first_div"] < r.get("horizon", 0)]
    dit = np.array([abs(r["iters_gpu"] - r["iters_cpu"]) for r in rows if "iters_cpu" in r])
    hist = {str(int(v)): int(c) for v, c in zip(*np.unique(dit, return_counts=True))} if dit.size else {}
    summary = dict(
        cfg=cfg, instances=count, linear=linear, gpu_wall_s=wall, cpu_wall_s=cpu_wall, cpu_workers=min(workers, count),
        cpu_seconds_sum=float(sum(r.get("cpu_s", 0.0) for r in rows)),
        identical=ident, status_equal=sum(r.get("status_gpu") == r.get("status_cpu", r.get("status_gpu")) for r in rows),
        never_hit_horizon=len(never), never_hit_horizon_identical=sum(r["identical"] for r in never),
        hit_horizon=len(post), hit_horizon_identical=sum(r["identical"] for r in post),
        post_horizon_share=len(post) / max(1, count),
        diverged_before_horizon=len(strict_fail),
        abs_iter_diff_hist=hist,
        max_x_rel_identical=max([r["x_rel"] for r in rows if r["identical"] and "x_rel" in r] or [0.0]),
        max_x_rel_all=max([r["x_rel"] for r in rows if "x_rel" in r] or [0.0]),
        max_lamb_rel_before_div=max([r.get("lamb_rel_before_div", 0.0) for r in rows] or [0.0]),
        median_lamb_rel_before_div=float(np.median([r.get("lamb_rel_before_div", 0.0) for r in rows])),
        iterations_cpu=dict(min=int(min(r["iters_cpu"] for r in rows if "iters_cpu" in r)),
                            max=int(max(r["iters_cpu"] for r in rows if "iters_cpu" in r))),
    )
    mism = [r for r in rows if not r["identical"]]
    return dict(summary=summary, mismatches=mism[:64])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", default="2,3,4")
    ap.add_argument("--count", default="128,128,8")
    ap.add_argument("--stages", type=int, default=128)
    ap.add_argument("--iteration-limit", type=int, default=None)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    cfgs = [int(c) for c in args.cfg.split(",")]
    counts = [int(c) for c in args.count.split(",")]
    import torch

    out = dict(device=torch.cuda.get_device_name(0), cpu_count=os.cpu_count(), theta_noise=THETA_NOISE,
               lamb_rtol=LAMB_RTOL, results=[])
    for cfg, cnt in zip(cfgs, counts):
        r = sweep(cfg, cnt, stages=args.stages, limit=args.iteration_limit)
        print(json.dumps(r["summary"]), flush=True)
        out["results"].append(r)
    if args.out:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        with open(args.out, "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
