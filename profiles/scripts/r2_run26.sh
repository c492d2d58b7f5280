set -x
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 200 python profiles/microbench/kron_determinism.py $TT 8 6 noref > gpurun_out/r2G_$name.log 2>&1; echo "$name rc=$?"; grep -v Warn gpurun_out/r2G_$name.log | tail -7; }
TT=12288
run tma1_s4 NMGP_KRON_SLOTS=4
run tma0_s4 NMGP_KRON_SLOTS=4 NMGP_GEMM_TMA=0
run tma1_s4_pb256 NMGP_KRON_SLOTS=4 NMGP_POTRF_PB=256
run tma1_s2 NMGP_KRON_SLOTS=2
TT=8192
run T8192_tma1_s4 NMGP_KRON_SLOTS=4
