import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from pygradflow_b200 import kernels as K
def run(N, seed=0):
    rng = np.random.default_rng(seed)
    M = rng.standard_normal((N, N)) + 0.1 * np.eye(N)
    r = rng.standard_normal(N)
    Kt = torch.as_tensor(M, device='cuda')[None].contiguous().clone()
    rhs = torch.as_tensor(r, device='cuda')[None].clone()
    piv = torch.zeros((1, N), dtype=torch.int32, device='cuda'); info = torch.zeros((1,), dtype=torch.int32, device='cuda')
    Nv = torch.full((1,), N, dtype=torch.int32, device='cuda')
    w = K.WorkList.all(1)
    K.lu_factor(Kt, N, Nv, piv, info, w)
    K.lu_solve(Kt, N, Nv, piv, rhs, False, w)
    x = rhs[0].cpu().numpy()
    import scipy.linalg
    _, lp = scipy.linalg.lu_factor(M.T)
    return float(np.abs(M @ x - r).max()), bool(np.array_equal(piv[0].cpu().numpy(), lp)), int(np.argmax(piv[0].cpu().numpy() != lp)) if not np.array_equal(piv[0].cpu().numpy(), lp) else -1
for N in (113, 128, 130, 160, 200, 300):
    print(N, run(N))
