import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, torch
from oracle import gradflow_oracle as orc
from pygradflow_b200 import synth
from pygradflow_b200.params import Params
from pygradflow_b200.problem import BatchedQP
from pygradflow_b200.solver import BatchedSolver
from test_gpu_round2 import _nonconvex_qp_batch
B,n,m=8,12,4
d=_nonconvex_qp_batch(B,n,m)
prob = BatchedQP(d["H"], d["A"], d["g"], d["b"], d["lb"], d["ub"])
s=BatchedSolver(prob, Params(inertia_correction=True, iteration_limit=14))
rows=[]
def hook(outer, sv):
    e=sv.engine
    rows.append((outer,int(sv.phase[0]),float(sv.lamb[0]),float(sv.lamb_next[0]),int(e.nneg[0]),int(e.info[0]),int(e.fbkey[0]),int(e.nI[0]), sv.mid[0][0].cpu().numpy().copy(), sv.fin[0][0].cpu().numpy().copy(), float(sv.theta[0])))
s.solve(d["x0"],d["y0"],on_iteration=hook)
p = orc.DenseQP(d["H"][0], d["A"][0], d["g"][0], d["b"][0], d["lb"][0], d["ub"][0])
ref = orc.Solver(p, orc.OracleParams(inertia_correction=True, iteration_limit=14, linear_solver="lapack")).solve(d["x0"][0], d["y0"][0], record=True)
for r,t in zip(rows,ref.trace):
    xg = r[8] if r[1]==2 else r[9]
    print(r[:8], "theta",r[10],"| cpu steps",t["newton_steps"],"acc",t["accept"],"lamb",t["lamb_used"],t["lamb_next"],"theta",t["theta"],"nA",None if t["active"] is None else int(t["active"].sum()), "xdiff", float(np.max(np.abs(xg-t["x"]))))
