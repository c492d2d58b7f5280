import sys, torch
sys.path.insert(0, "/root/repo")
from collaborative_nonstationary_multivariate_gaussian_process_b200 import _ops as ops
T = int(sys.argv[1])
g = torch.Generator().manual_seed(0)
A = torch.randn(T, T, generator=g, dtype=torch.float64)
A = (A @ A.t() / T + torch.eye(T, dtype=torch.float64)).cuda()
w = A.clone()
ops.potrf_big(w)
torch.cuda.synchronize()
w.copy_(A)
torch.cuda.synchronize()
ops.potrf_big(w)
torch.cuda.synchronize()
