set -x
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_kernels_gpu.py tests/test_simcode_gpu.py tests/test_dsvi_gpu.py -x -q > gpurun_out/r2x_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2x_pytest.log
tail -3 gpurun_out/r2x_pytest.log
B="--cpu-baseline skip --no-e2e --others skip"
timeout 200 python bench.py --workload ecog --steps 5 --warmup 3 $B > gpurun_out/r2x_ecog.json 2> gpurun_out/r2x_ecog.err; echo "ecog rc=$?"
timeout 200 python bench.py --workload pm25 --steps 10 --warmup 3 $B > gpurun_out/r2x_pm25.json 2> gpurun_out/r2x_pm25.err; echo "pm25 rc=$?"
timeout 200 python bench.py --workload hcp --steps 10 --warmup 3 $B > gpurun_out/r2x_hcp.json 2> gpurun_out/r2x_hcp.err; echo "hcp rc=$?"
timeout 300 python bench.py --workload sweep --sweep-T 8192 --sweep-D 128 --steps 1 --warmup 1 --cpu-baseline skip > gpurun_out/r2x_sweep.json 2> gpurun_out/r2x_sweep.err; echo "sweep rc=$?"
timeout 100 python - > gpurun_out/r2x_eigh.log 2>&1 <<'PY'
import torch, time
from collaborative_nonstationary_multivariate_gaussian_process_b200 import _ops as ops
for n in (15, 64, 128):
    A = torch.randn(n, n, dtype=torch.float64, device="cuda"); A = A @ A.T + torch.eye(n, dtype=torch.float64, device="cuda")
    ops.eigh_small(A); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); w, V = ops.eigh_small(A); e1.record(); torch.cuda.synchronize()
    wr = torch.linalg.eigvalsh(A)
    print(n, "ms", e0.elapsed_time(e1), "eig err", float((w - wr).abs().max() / wr.abs().max()), "recon", float((V @ torch.diag(w) @ V.T - A).abs().max() / A.abs().max()))
PY
cat gpurun_out/r2x_eigh.log
