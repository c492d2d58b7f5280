set -x
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 60 python profiles/microbench/kron_determinism.py 12288 8 5 noref > gpurun_out/r2K_$name.log 2>&1; echo "$name rc=$?"; grep -v Warn gpurun_out/r2K_$name.log | tail -6; }
run tma_noprefetch NMGP_KRON_SLOTS=4 NMGP_TMA_CONCURRENT=1 NMGP_TMA_NOPREFETCH=1
run tma_promo0 NMGP_KRON_SLOTS=4 NMGP_TMA_CONCURRENT=1 NMGP_TMA_L2PROMO=0
run tma_plain NMGP_KRON_SLOTS=4 NMGP_TMA_CONCURRENT=1
