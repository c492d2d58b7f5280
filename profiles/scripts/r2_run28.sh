set -x
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -x -q > gpurun_out/r2I_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2I_pytest.log
tail -3 gpurun_out/r2I_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2I_smoke.log 2>&1; echo "smoke rc=$?"
for T in 12288 16384; do
timeout 400 python bench.py --workload sweep --sweep-T $T --sweep-D 128 --steps 1 --warmup 1 --cpu-baseline skip > gpurun_out/r2I_sweep_${T}x128_1gpu.json 2> gpurun_out/r2I_sweep_${T}.err; echo "sweep $T rc=$?"
done
timeout 500 python bench.py --steps 5 --warmup 3 > gpurun_out/r2I_bench_default.json 2> gpurun_out/r2I_bench_default.err; echo "bench rc=$?"
tail -c 300 gpurun_out/r2I_bench_default.err
