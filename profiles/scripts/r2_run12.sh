set -x
mkdir -p gpurun_out
B="--cpu-baseline skip --no-e2e --no-graph"
timeout 600 python -m pytest tests/test_sim_logpos_gpu.py tests/test_simcode_gpu.py -x -q > gpurun_out/r2q_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2q_pytest.log
tail -5 gpurun_out/r2q_pytest.log
timeout 300 python bench.py --workload ecog --steps 1 --warmup 1 $B > gpurun_out/r2q_plain_ecog.json 2> gpurun_out/r2q_plain_ecog.err; echo "ecog rc=$?"
timeout 300 python bench.py --workload pm25 --steps 1 --warmup 1 $B > gpurun_out/r2q_plain_pm25.json 2> gpurun_out/r2q_plain_pm25.err; echo "pm25 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2q_launches_ecog.csv python bench.py --workload ecog --steps 1 --warmup 1 $B > gpurun_out/r2q_ncu_launch_ecog.log 2>&1; echo "launchlist rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:"k_latent_fused|k_gram_mma" -s 17 -c 2 -o gpurun_out/r2q_prof_ecog python bench.py --workload ecog --steps 1 --warmup 1 $B > gpurun_out/r2q_ncu_ecog.log 2>&1; echo "ncu ecog rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:"k_lq|k_gram_mma" -s 12 -c 3 -o gpurun_out/r2q_prof_pm25 python bench.py --workload pm25 --steps 1 --warmup 1 $B > gpurun_out/r2q_ncu_pm25.log 2>&1; echo "ncu pm25 rc=$?"
ls -la gpurun_out/; du -sh gpurun_out
