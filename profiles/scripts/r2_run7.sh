set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2g_pytest.log
tail -5 gpurun_out/r2g_pytest.log
for ns in 2 4 6 8; do
NMGP_KRON_SLOTS=$ns timeout 600 python bench.py --workload sweep --sweep-T 8192 --sweep-D 128 --steps 1 --warmup 1 --cpu-baseline skip > gpurun_out/r2g_sweep_8192x128_s$ns.json 2> gpurun_out/r2g_sweep_s$ns.err; echo "rc=$?"
done
NMGP_KRON_SLOTS=4 timeout 600 python bench.py --workload sweep --sweep-T 16384 --sweep-D 16 --steps 1 --warmup 1 --cpu-baseline skip > gpurun_out/r2g_sweep_16384x16.json 2> gpurun_out/r2g_sweep_16384x16.err; echo "rc=$?"
NMGP_KRON_SLOTS=4 timeout 600 python bench.py --workload sweep --sweep-T 4096 --sweep-D 64 --steps 2 --warmup 1 --cpu-baseline skip > gpurun_out/r2g_sweep_4096x64.json 2> gpurun_out/r2g_sweep_4096x64.err; echo "rc=$?"
tail -n 3 gpurun_out/r2g_sweep_s4.err
