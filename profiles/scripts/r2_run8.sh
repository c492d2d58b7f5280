set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h_pytest.log
tail -5 gpurun_out/r2h_pytest.log
for pb in 512 1024; do
NMGP_POTRF_PB=$pb NMGP_KRON_SLOTS=4 timeout 600 python bench.py --workload sweep --sweep-T 8192 --sweep-D 128 --steps 1 --warmup 1 --cpu-baseline skip > gpurun_out/r2h_sweep_8192x128_pb$pb.json 2> gpurun_out/r2h_sweep_pb$pb.err; echo "rc=$?"
done
NMGP_POTRF_PB=512 NMGP_KRON_SLOTS=4 timeout 600 python bench.py --workload sweep --sweep-T 4096 --sweep-D 64 --steps 2 --warmup 1 --cpu-baseline skip > gpurun_out/r2h_sweep_4096x64_pb512.json 2> gpurun_out/r2h_sweep_4096_pb512.err; echo "rc=$?"
NMGP_POTRF_PB=256 NMGP_KRON_SLOTS=4 timeout 600 python bench.py --workload sweep --sweep-T 4096 --sweep-D 64 --steps 2 --warmup 1 --cpu-baseline skip > gpurun_out/r2h_sweep_4096x64_pb256.json 2> gpurun_out/r2h_sweep_4096_pb256.err; echo "rc=$?"
NMGP_POTRF_PB=1024 NMGP_KRON_SLOTS=4 timeout 600 python bench.py --workload sweep --sweep-T 16384 --sweep-D 16 --steps 1 --warmup 1 --cpu-baseline skip > gpurun_out/r2h_sweep_16384x16_pb1024.json 2> gpurun_out/r2h_sweep_16384_pb1024.err; echo "rc=$?"
# ncu: tensor-pipe utilisation of the GEMM / Cholesky kernels (one blocked Cholesky + one GEMM at T=8192)
timeout 300 python bench_sweep.py --sizes 8192 > gpurun_out/r2h_sweep8k_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_gemm_nt|k_potrf_diag_inv" -s 40 -c 12 -o gpurun_out/r2h_prof_chol python bench_sweep.py --sizes 8192 > gpurun_out/r2h_ncu_chol.log 2>&1
echo ncu rc=$?
