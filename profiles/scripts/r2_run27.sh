set -x
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 300 python profiles/microbench/kron_determinism.py $TT 8 $REPS noref > gpurun_out/r2H_$name.log 2>&1; echo "$name rc=$?"; grep -v Warn gpurun_out/r2H_$name.log | tail -9; }
TT=12288; REPS=8
run fix_s4 NMGP_KRON_SLOTS=4
run single_s1_tma NMGP_KRON_SLOTS=1
TT=16384; REPS=4
run fix_s4_T16384 NMGP_KRON_SLOTS=4
timeout 300 python bench.py --workload sweep --sweep-T 8192 --sweep-D 128 --steps 2 --warmup 1 --cpu-baseline skip > gpurun_out/r2H_sweep.json 2> gpurun_out/r2H_sweep.err; echo "sweep rc=$?"
