set -x
N=${1:-8}
mkdir -p gpurun_out
run() { name=$1; shift
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "$name rc=$?"
  tail -n 2 gpurun_out/$name.err
}
run r2E_default_${N}gpu --steps 10 --warmup 3 --cpu-baseline skip
run r2E_sweep16384_${N}gpu --workload sweep --sweep-T 16384 --sweep-D 128 --steps 1 --warmup 1 --cpu-baseline skip
run r2E_sweep12288_${N}gpu --workload sweep --sweep-T 12288 --sweep-D 128 --steps 1 --warmup 1 --cpu-baseline skip
