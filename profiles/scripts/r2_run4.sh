set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
tail -5 gpurun_out/r2d_pytest.log
timeout 600 python bench_sweep.py --sizes 4096,8192,16384 > gpurun_out/r2d_sweep.jsonl 2> gpurun_out/r2d_sweep.err; echo rc=$?
tail -n 3 gpurun_out/r2d_sweep.err
timeout 300 python bench_sweep.py --sizes 16384 > gpurun_out/r2d_sweep16k_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_simcov -c 2 -o gpurun_out/r2d_prof_simcov python bench_sweep.py --sizes 16384 > gpurun_out/r2d_ncu_simcov.log 2>&1
echo ncu rc=$?
timeout 300 python bench.py --workload pm25 --steps 1 --warmup 3 --cpu-baseline skip --no-e2e > gpurun_out/r2d_pm25_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2d_launches_pm25.csv python bench.py --workload pm25 --steps 1 --warmup 3 --cpu-baseline skip --no-e2e > gpurun_out/r2d_ncu_pm25.log 2>&1
echo ncu rc=$?
