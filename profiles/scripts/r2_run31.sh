set -x
mkdir -p gpurun_out
timeout 60 python bench_sweep.py --sizes 8192 > gpurun_out/r2M_sweep8k_plain.log 2>&1; echo "plain rc=$?"
timeout 100 ncu --set full --clock-control none -k regex:k_gemm_nt_tma -s 30 -c 2 -o gpurun_out/r2M_prof_gemm_tma python bench_sweep.py --sizes 8192 > gpurun_out/r2M_ncu.log 2>&1; echo "ncu rc=$?"
