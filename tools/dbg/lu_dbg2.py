import sys, os, numpy as np, torch
sys.path.insert(0, '/root/repo')
from pygradflow_b200 import kernels as K
N = int(sys.argv[1]); tag = sys.argv[2]
rng = np.random.default_rng(0)
M = rng.standard_normal((N, N)) + 0.1 * np.eye(N)
Kt = torch.as_tensor(M, device='cuda')[None].contiguous().clone()
piv = torch.zeros((1, N), dtype=torch.int32, device='cuda'); info = torch.zeros((1,), dtype=torch.int32, device='cuda')
Nv = torch.full((1,), N, dtype=torch.int32, device='cuda')
K.lu_factor(Kt, N, Nv, piv, info, K.WorkList.all(1))
np.save(f'/tmp/lu_{tag}_{N}.npy', Kt[0].cpu().numpy()); np.save(f'/tmp/piv_{tag}_{N}.npy', piv[0].cpu().numpy())
if tag == 'delayed':
    A = np.load(f'/tmp/lu_nodelay_{N}.npy'); B = Kt[0].cpu().numpy()
    D = np.abs(A - B) > 1e-9 * (1 + np.abs(A))
    rows = np.where(D.any(axis=1))[0]; cols = np.where(D.any(axis=0))[0]
    print(N, 'diff rows', rows[:10], '... count', len(rows), 'cols', cols[:10], 'count', len(cols))
    # block summary 32x32
    nb = (N + 31) // 32
    for bi in range(nb):
        print(''.join('X' if D[bi*32:(bi+1)*32, bj*32:(bj+1)*32].any() else '.' for bj in range(nb)))
