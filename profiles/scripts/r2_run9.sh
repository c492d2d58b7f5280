set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_simcode_gpu.py -x -q > gpurun_out/r2i_pytest_sim.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2i_pytest_sim.log
tail -5 gpurun_out/r2i_pytest_sim.log
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_simcode_gpu.py > gpurun_out/r2i_pytest_rest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2i_pytest_rest.log
tail -3 gpurun_out/r2i_pytest_rest.log
timeout 600 python bench_sweep.py --sizes 4096,8192,16384 > gpurun_out/r2i_sweep_tma.jsonl 2> gpurun_out/r2i_sweep_tma.err; echo rc=$?
NMGP_GEMM_TMA=0 timeout 600 python bench_sweep.py --sizes 8192 > gpurun_out/r2i_sweep_notma.jsonl 2> gpurun_out/r2i_sweep_notma.err; echo rc=$?
timeout 600 python bench.py --workload sweep --sweep-T 8192 --sweep-D 128 --steps 1 --warmup 1 --cpu-baseline skip > gpurun_out/r2i_sweep_8192x128.json 2> gpurun_out/r2i_sweep_8192x128.err; echo "rc=$?"
tail -n 3 gpurun_out/r2i_sweep_8192x128.err
