set -x
mkdir -p gpurun_out
timeout 150 python bench.py --steps 3 --warmup 3 --cpu-baseline skip > gpurun_out/r2L_bench.json 2> gpurun_out/r2L_bench.err; echo "bench rc=$?"
tail -c 300 gpurun_out/r2L_bench.err
