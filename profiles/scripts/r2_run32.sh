set -x
mkdir -p gpurun_out
timeout 80 python -m pytest tests/test_kernels_gpu.py -x -q -k "latent_fused or gram" > gpurun_out/r2N_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2N_pytest.log
tail -2 gpurun_out/r2N_pytest.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2N_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2N_smoke.log
