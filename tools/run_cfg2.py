"""cfg2 at full size: 4096 chained-Rosenbrock instances (n = 64, bounds only) through the fused persistent solver
(gf_rosen_fused_solve) and, for comparison, through the lock-step driver; a sample of instances against the CPU oracle.

    python tools/run_cfg2.py [--batch 4096] [--oracle 128] [--lockstep] [--out profiles/r02_cfg2_fused.json]
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


def _oracle(k):
    from threadpoolctl import threadpool_limits

    threadpool_limits(1)
    from oracle import gradflow_oracle as orc
    from pygradflow_b200 import synth

    d = synth.rosenbrock_instance(k, 64)
    p = orc.ChainedRosenbrock(d["a"], d["b"], d["lb"], d["ub"])
    t0 = time.perf_counter()
    r = orc.Solver(p, orc.OracleParams()).solve(d["x0"], d["y0"], record=True)
    th = [t["theta"] for t in r.trace]
    horizon = next((i for i, v in enumerate(th) if v == v and v < 1e-8), len(th))
    return dict(k=k, status=int(r.status), iterations=int(r.iterations), accepted=int(r.accepted_steps), x=r.x,
                pre_horizon=horizon >= len(th), cpu_s=time.perf_counter() - t0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--oracle", type=int, default=128)
    ap.add_argument("--lockstep", action="store_true")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    import torch

    from pygradflow_b200 import synth
    from pygradflow_b200.params import Params
    from pygradflow_b200.problem import BatchedRosenbrock
    from pygradflow_b200.solver import BatchedSolver

    d = synth.rosenbrock_batch(range(a.batch), 64)
    prob = BatchedRosenbrock(d["a"], d["b"], d["lb"], d["ub"])
    out = dict(config="cfg2: chained Rosenbrock n=64, bounds only, default Params", batch=a.batch)
    for name, fused in (("fused", True),) + ((("lockstep", False),) if a.lockstep else ()):
        s = BatchedSolver(prob, Params(fused=fused))
        s.solve(d["x0"], None)  # warm-up (module load, allocator)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = s.solve(d["x0"], None)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        it = res.iterations.cpu().numpy()
        out[name] = dict(wall_s=wall, solves_per_s=a.batch / wall, status_counts=np.bincount(res.status.cpu().numpy()).tolist(),
                         iterations_median=float(np.median(it)), iterations_max=int(it.max()),
                         iterations_total=int(it.sum()), newton_steps=res.newton_steps,
                         launches=getattr(s, "fused_launches", None),
                         us_per_outer_iteration_of_the_slowest_instance=1e6 * wall / max(1, int(it.max())),
                         us_per_outer_iteration_aggregate=1e6 * wall / max(1, int(it.sum())))
        print(name, json.dumps(out[name]), flush=True)
        if fused:
            fres = res
    if a.oracle > 0:
        ks = list(range(a.oracle))
        with ProcessPoolExecutor(os.cpu_count()) as ex:
            refs = list(ex.map(_oracle, ks))
        x = fres.x.cpu().numpy()
        it = fres.iterations.cpu().numpy()
        acc = fres.accepted_steps.cpu().numpy()
        st = fres.status.cpu().numpy()
        pre = [r for r in refs if r["pre_horizon"]]
        same = sum(1 for r in pre if (st[r["k"]], it[r["k"]], acc[r["k"]]) == (r["status"], r["iterations"], r["accepted"]))
        xerr = max(float(np.max(np.abs(x[r["k"]] - r["x"])) / max(1.0, float(np.max(np.abs(r["x"]))))) for r in pre)
        post = [r for r in refs if not r["pre_horizon"]]
        out["parity_vs_oracle"] = dict(instances=len(refs), pre_horizon=len(pre),
                                       pre_horizon_identical_status_iterations_accepted=same, pre_horizon_max_rel_err_x=xerr,
                                       post_horizon=len(post),
                                       post_horizon_same_status=sum(1 for r in post if st[r["k"]] == r["status"]),
                                       oracle_cpu_s_per_instance_mean=float(np.mean([r["cpu_s"] for r in refs])),
                                       mismatches=[dict(k=r["k"], gpu=[int(st[r["k"]]), int(it[r["k"]]), int(acc[r["k"]])],
                                                        cpu=[r["status"], r["iterations"], r["accepted"]]) for r in pre
                                                   if (st[r["k"]], it[r["k"]], acc[r["k"]]) != (r["status"], r["iterations"], r["accepted"])][:10])
        print("parity", json.dumps(out["parity_vs_oracle"]), flush=True)
    if a.out:
        with open(a.out, "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
