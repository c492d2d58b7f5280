set -x
mkdir -p gpurun_out
run() { name=$1; n=$2; shift 2
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "$name rc=$?"
  tail -n 2 gpurun_out/$name.err
}
run r2o_ecog_8gpu_graph 8 --steps 20 --warmup 5 --cpu-baseline skip
run r2o_hcp_8gpu_graph 8 --workload hcp --steps 20 --warmup 5 --cpu-baseline skip
run r2o_sweep_8gpu 8 --workload sweep --sweep-T 8192 --sweep-D 128 --steps 2 --warmup 1 --cpu-baseline skip
run r2o_ecog_4gpu_graph 4 --steps 20 --warmup 5 --cpu-baseline skip
