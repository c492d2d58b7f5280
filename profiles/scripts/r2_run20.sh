set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/r2z_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2z_pytest.log
tail -3 gpurun_out/r2z_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2z_smoke.log 2>&1; echo "smoke rc=$?"
tail -2 gpurun_out/r2z_smoke.log
timeout 100 python - > gpurun_out/r2z_eigh.log 2>&1 <<'PY'
import torch
from collaborative_nonstationary_multivariate_gaussian_process_b200 import _ops as ops
for n in (15, 64, 128):
    A = torch.randn(n, n, dtype=torch.float64, device="cuda"); A = A @ A.T + torch.eye(n, dtype=torch.float64, device="cuda")
    ops.eigh_small(A); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); w, V = ops.eigh_small(A); e1.record(); torch.cuda.synchronize()
    wr = torch.linalg.eigvalsh(A)
    print(n, "ms", e0.elapsed_time(e1), "eig err", float((w - wr).abs().max() / wr.abs().max()), "recon", float((V @ torch.diag(w) @ V.T - A).abs().max() / A.abs().max()), "orth", float((V.T @ V - torch.eye(n, dtype=torch.float64, device="cuda")).abs().max()))
PY
cat gpurun_out/r2z_eigh.log
timeout 300 python bench.py --workload sweep --sweep-T 8192 --sweep-D 128 --steps 2 --warmup 1 --cpu-baseline skip > gpurun_out/r2z_sweep.json 2> gpurun_out/r2z_sweep.err; echo "sweep rc=$?"
