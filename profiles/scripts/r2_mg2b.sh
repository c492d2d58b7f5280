set -x
mkdir -p gpurun_out
run() { name=$1; n=$2; shift 2
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "$name rc=$?"
  tail -n 2 gpurun_out/$name.err
}
run r2n_tiny_2gpu_graph 2 --workload tiny --steps 10 --warmup 3 --cpu-baseline skip
run r2n_hcp_2gpu_graph 2 --workload hcp --steps 10 --warmup 3 --cpu-baseline skip
