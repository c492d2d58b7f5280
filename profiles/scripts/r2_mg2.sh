set -x
mkdir -p gpurun_out
run() { # name ngpu args...
  name=$1; n=$2; shift 2
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "$name rc=$?"
  tail -n 4 gpurun_out/$name.err
}
run r2m_ecog_2gpu_graph 2 --steps 10 --warmup 3 --cpu-baseline skip
run r2m_ecog_2gpu_nograph 2 --steps 10 --warmup 3 --cpu-baseline skip --no-graph
run r2m_hcp_2gpu_graph 2 --workload hcp --steps 10 --warmup 3 --cpu-baseline skip
run r2m_sweep_2gpu 2 --workload sweep --sweep-T 8192 --sweep-D 128 --steps 1 --warmup 1 --cpu-baseline skip
