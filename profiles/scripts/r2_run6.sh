set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log
tail -5 gpurun_out/r2f_pytest.log
timeout 600 python bench.py --workload sweep --sweep-T 8192 --sweep-D 128 --steps 2 --warmup 1 > gpurun_out/r2f_sweep_8192x128.json 2> gpurun_out/r2f_sweep_8192x128.err; echo "rc=$?"
tail -n 3 gpurun_out/r2f_sweep_8192x128.err
timeout 600 python bench.py --workload sweep --sweep-T 16384 --sweep-D 16 --steps 1 --warmup 1 --cpu-baseline skip > gpurun_out/r2f_sweep_16384x16.json 2> gpurun_out/r2f_sweep_16384x16.err; echo "rc=$?"
tail -n 3 gpurun_out/r2f_sweep_16384x16.err
timeout 600 python bench.py --workload sweep --sweep-T 4096 --sweep-D 32 --steps 3 --warmup 1 --cpu-baseline skip > gpurun_out/r2f_sweep_4096x32.json 2> gpurun_out/r2f_sweep_4096x32.err; echo "rc=$?"
timeout 300 python bench.py --impl reference --workload sweep --steps 3 > gpurun_out/r2f_ref_sweep.json 2>&1; echo "rc=$?"
