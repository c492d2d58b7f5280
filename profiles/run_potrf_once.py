"""Runs one blocked Cholesky (T from argv, default 4096) -- used under ncu to list per-kernel times."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from collaborative_nonstationary_multivariate_gaussian_process_b200 import _ops as ops
from oracle import nmgp_oracle as orc
T = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
g = torch.Generator().manual_seed(T)
x = torch.sort(torch.rand(T, generator=g, dtype=torch.float64))[0].view(-1, 1)
K = (orc.sim_nonstationary_cov(x, ell1=torch.exp(3 * (x.view(-1) - 1) ** 3 - 3.0)) + 1e-2 * torch.eye(T, dtype=torch.float64)).cuda()
for _ in range(2):
    A = K.clone()
    ops.potrf_big(A)
torch.cuda.synchronize()
print("ok", float(A[5, 5]))
