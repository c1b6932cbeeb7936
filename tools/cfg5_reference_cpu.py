"""cfg5, CPU column: the REAL reference's linear solver (pygradflow.linear_solver.lu_solver.LUSolver, imported from
/root/reference -- build container only, it cannot travel to the GPU box) on the cfg5 matrices of
pygradflow_b200.synth.kkt_instance: factorisation (constructor) and one solve, best of `--reps`, one core.

    python tools/cfg5_reference_cpu.py [--out profiles/r02_cfg5_reference_lusolver.json]

tools/sweep_cfg5.py merges the file into its tables (column "reference LUSolver, build container")."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden", "_stubs"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)
import numpy as np
import scipy.sparse as sps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r02_cfg5_reference_lusolver.json"))
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--ns", default="32,64,128,256,512,1024,2048")
    a = ap.parse_args()
    from threadpoolctl import threadpool_limits

    threadpool_limits(1)
    from pygradflow.linear_solver.lu_solver import LUSolver

    from pygradflow_b200 import synth

    rows = {}
    for N in [int(v) for v in a.ns.split(",")]:
        K, rhs, m = synth.kkt_instance(N)
        A = sps.csc_matrix(K)
        tf, ts = [], []
        for _ in range(a.reps):
            t0 = time.perf_counter()
            s = LUSolver(A, symmetric=True)
            t1 = time.perf_counter()
            x = s.solve(rhs)
            t2 = time.perf_counter()
            tf.append(t1 - t0)
            ts.append(t2 - t1)
        res = float(np.max(np.abs(K @ x - rhs)))
        rows[str(N)] = dict(N=N, factor_ms=1e3 * min(tf), solve_ms=1e3 * min(ts), residual=res)
        print(N, rows[str(N)], flush=True)
    import platform

    meta = dict(what="pygradflow.linear_solver.lu_solver.LUSolver (scipy splu) on synth.kkt_instance(N), one core, best of %d" % a.reps,
                where="build container", cpu=platform.processor() or platform.machine(), cores=os.cpu_count())
    with open(a.out, "w") as f:
        json.dump(dict(meta=meta, rows=rows), f, indent=1)


if __name__ == "__main__":
    main()
