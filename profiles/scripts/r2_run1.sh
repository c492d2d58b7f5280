set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench_ecog.json 2> gpurun_out/r2a_bench_ecog.err; echo rc=$?
python bench.py --workload pm25 --steps 10 --warmup 3 --cpu-baseline skip > gpurun_out/r2a_bench_pm25.json 2> gpurun_out/r2a_bench_pm25.err; echo rc=$?
python bench.py --workload hcp --steps 5 --warmup 3 --cpu-baseline skip > gpurun_out/r2a_bench_hcp.json 2> gpurun_out/r2a_bench_hcp.err; echo rc=$?
python bench.py --workload sim --steps 50 --warmup 5 > gpurun_out/r2a_bench_sim.json 2> gpurun_out/r2a_bench_sim.err; echo rc=$?
python bench.py --impl reference --workload sim --steps 50 --warmup 5 > gpurun_out/r2a_ref_sim.json 2>&1; echo rc=$?
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2a_ref_ecog.json 2> gpurun_out/r2a_ref_ecog.err; echo rc=$?
nproc; lscpu | grep "Model name"
tail -c 600 gpurun_out/r2a_pytest.log
