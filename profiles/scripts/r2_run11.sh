set -x
mkdir -p gpurun_out
B="--cpu-baseline skip --no-e2e --no-graph"
# plain runs first (must exit 0 without ncu)
timeout 300 python bench.py --workload ecog --steps 1 --warmup 1 $B > gpurun_out/r2p_plain_ecog.json 2> gpurun_out/r2p_plain_ecog.err; echo "ecog rc=$?"
timeout 300 python bench.py --workload pm25 --steps 1 --warmup 1 $B > gpurun_out/r2p_plain_pm25.json 2> gpurun_out/r2p_plain_pm25.err; echo "pm25 rc=$?"
# launch list of the ecog step (eager, one warm-up + one timed step)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2p_launches_ecog.csv python bench.py --workload ecog --steps 1 --warmup 1 $B > gpurun_out/r2p_ncu_launch_ecog.log 2>&1; echo "launchlist rc=$?"
# full captures of the dominant kernels
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_latent_fused|k_gram_mma" -s 17 -c 4 -o gpurun_out/r2p_prof_ecog python bench.py --workload ecog --steps 1 --warmup 1 $B > gpurun_out/r2p_ncu_ecog.log 2>&1; echo "ncu ecog rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_lq|k_gram_mma|k_solve_rows" -s 20 -c 8 -o gpurun_out/r2p_prof_pm25 python bench.py --workload pm25 --steps 1 --warmup 1 $B > gpurun_out/r2p_ncu_pm25.log 2>&1; echo "ncu pm25 rc=$?"
ls -la gpurun_out/r2p*
