import sys, ctypes, torch
sys.path.insert(0, "/root/repo")
from collaborative_nonstationary_multivariate_gaussian_process_b200 import _ops as ops, _lib
g = torch.Generator().manual_seed(0)
A = torch.randn(128, 128, generator=g, dtype=torch.float64); A = (A @ A.t() / 128 + torch.eye(128, dtype=torch.float64)).cuda()
for _ in range(3):
    w = A.clone(); ops.potrf_big(w); torch.cuda.synchronize()
buf = (ctypes.c_longlong * 8)()
lib = _lib.lib()
print(lib.nmgp_debug_clk(buf), list(buf), sum(buf))
print("labels: load, factor16(x8), dinv(x8), P2(x8), P3(x8), store, linv")
