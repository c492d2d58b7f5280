"""Determinism / accuracy probe of the concurrent eigen-block pipeline at large T: per-block half log-determinants and
quadratic forms from kronecker_operation.block_pipeline (NMGP_KRON_SLOTS slots) against torch.linalg.cholesky."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench_sweep  # noqa: E402
from collaborative_nonstationary_multivariate_gaussian_process_b200 import kernels, kronecker_operation as ko  # noqa: E402
from collaborative_nonstationary_multivariate_gaussian_process_b200 import _ops as ops  # noqa: E402

T, D = int(sys.argv[1]), int(sys.argv[2])
REPS = int(sys.argv[3]) if len(sys.argv) > 3 else 3
REF = (sys.argv[4] != "noref") if len(sys.argv) > 4 else True
dev = torch.device("cuda", 0)
x, ell, Bf, y, s2 = bench_sweep.sweep_problem(T, D, dev)
K = kernels.Nonstationary_RBF_cov(x, ell1=ell)
lam, V = ops.eigh_small(Bf.contiguous())
Rt = (V.t() @ y.view(D, T)).contiguous()
runs = []
for rep in range(REPS):
    res = ko.block_pipeline(s2, Bf, K, Rt=Rt, want_alpha=False)
    torch.cuda.synchronize()
    runs.append((res["hld"].clone(), res["quad"].clone(), res["info"].clone()))
print("slots", ko.NSLOT, "info", runs[0][2].tolist())
for a in runs[1:]:
    print("vs run 0:  hld", float((a[0] - runs[0][0]).abs().max()), " quad", float((a[1] - runs[0][1]).abs().max()),
          "blocks differing", (a[1] != runs[0][1]).nonzero().view(-1).tolist())
res2 = ko.block_pipeline(s2, Bf, K, Rt=Rt, want_alpha=True)
torch.cuda.synchronize()
print("augmented vs solve route: hld", float((res2["hld"] - runs[0][0]).abs().max()), " quad rel", float(((res2["quad"] - runs[0][1]).abs() / runs[0][1].abs()).max()))
# reference: torch / cuSOLVER Cholesky per block
if not REF:
    sys.exit(0)
hr, qr = [], []
for m in range(D):
    A = K * lam[m] + torch.eye(T, dtype=torch.float64, device=dev) * s2.to(dev)
    L = torch.linalg.cholesky(A)
    hr.append(torch.log(torch.diagonal(L)).sum())
    z = torch.linalg.solve_triangular(L, Rt[m].view(-1, 1), upper=False)
    qr.append((z * z).sum())
    del A, L
hr, qr = torch.stack(hr), torch.stack(qr)
print("vs torch: hld rel", float(((runs[0][0] - hr).abs() / hr.abs()).max()), " quad rel", float(((runs[0][1] - qr).abs() / qr.abs()).max()))
print("quad", [float(v) for v in runs[0][1][:4]], [float(v) for v in qr[:4]])
print("lam", [float(v) for v in lam[:4]], "cond-ish", float(lam.max() * T / s2))
