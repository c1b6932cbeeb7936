"""NCCL check of the sharded solve (torchrun, one rank per GPU): every rank solves its block, one all-gather at the
end; rank 0 also solves the whole batch alone and compares bit for bit.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/run_sharded.py [--B 64] [--n 128] [--m 64]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from pygradflow_b200 import synth
from pygradflow_b200.dist import solve_sharded
from pygradflow_b200.problem import BatchedQP
from pygradflow_b200.solver import BatchedSolver

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=64)
ap.add_argument("--n", type=int, default=128)
ap.add_argument("--m", type=int, default=64)
ap.add_argument("--cfg", type=int, default=3, help="3: dense QPs (n, m); 4: optimal-control problems (--stages)")
ap.add_argument("--stages", type=int, default=128)
ap.add_argument("--no-check", action="store_true", help="skip the single-rank reference solve on rank 0")
ap.add_argument("--reps", type=int, default=2, help="sharded solves; the last one is timed (the first warms up)")
ap.add_argument("--out", default=None)
args = ap.parse_args()
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local_rank = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local_rank)
dev = torch.device("cuda", local_rank)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
from pygradflow_b200.dist import shard_range
from pygradflow_b200.problem import BatchedOCP

if args.cfg == 4:
    lo_, hi_ = (0, args.B) if (rank == 0 and not args.no_check) else shard_range(args.B, rank, world)
    d = synth.ocp_batch(range(lo_, hi_), stages=args.stages)
    d = {k: v for k, v in d.items()}
    base = lo_

    def factory(lo, hi):
        sl = slice(lo - base, hi - base)
        return BatchedOCP(d["A"][sl], d["B"][sl], d["Q"][sl], d["R"][sl], d["xinit"][sl], d["umax"], d["h"], device=dev)

    nvars, ncons = args.stages * 16, args.stages * 8
    d["x0"] = __import__("numpy").zeros((args.B, nvars))
    d["y0"] = __import__("numpy").zeros((args.B, ncons))
else:
    d = synth.qp_batch(range(args.B), args.n, args.m)

    def factory(lo, hi):
        return BatchedQP(d["H"][lo:hi], d["A"][lo:hi], d["g"][lo:hi], d["b"][lo:hi], d["lb"][lo:hi], d["ub"][lo:hi], device=dev)


x0 = torch.as_tensor(d["x0"], device=dev)
y0 = torch.as_tensor(d["y0"], device=dev)
for _ in range(max(1, args.reps)):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    res = solve_sharded(args.B, factory, None, x0, y0)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
if world > 1:  # the slowest rank counts
    wt = torch.tensor([wall], dtype=torch.float64, device=dev)
    dist.all_reduce(wt, op=dist.ReduceOp.MAX)
    wall = float(wt.item())
if rank == 0:
    out = dict(world=world, cfg=args.cfg, B=args.B, wall_s=wall, solves_per_s=args.B / wall,
               backend=dist.get_backend() if world > 1 else None, optimal=int((res.status == 1).sum().item()),
               max_iterations=int(res.iterations.max().item()))
    if not args.no_check:
        full = BatchedSolver(factory(0, args.B)).solve(x0, y0)
        out.update(x_equal=bool(torch.equal(res.x, full.x)), y_equal=bool(torch.equal(res.y, full.y)),
                   status_equal=bool(torch.equal(res.status, full.status)),
                   iterations_equal=bool(torch.equal(res.iterations, full.iterations)))
    print(json.dumps(out))
    if args.out:
        json.dump(out, open(args.out, "w"), indent=1)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
