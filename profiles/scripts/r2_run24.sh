set -x
mkdir -p gpurun_out
for T in 12288 16384; do
timeout 400 python bench.py --workload sweep --sweep-T $T --sweep-D 128 --steps 1 --warmup 1 --cpu-baseline skip > gpurun_out/r2D_sweep_${T}x128_1gpu.json 2> gpurun_out/r2D_sweep_${T}.err; echo "sweep $T rc=$?"
tail -n 2 gpurun_out/r2D_sweep_${T}.err
done
