set -x
N=${1:-8}
mkdir -p gpurun_out
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 10 --warmup 3 --cpu-baseline skip > gpurun_out/r2y_default_${N}gpu.json 2> gpurun_out/r2y_default_${N}gpu.err; echo "rc=$?"
tail -n 3 gpurun_out/r2y_default_${N}gpu.err
wc -c gpurun_out/r2y_default_${N}gpu.json
