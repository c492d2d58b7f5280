set -x
mkdir -p gpurun_out
B="--cpu-baseline skip --no-e2e --no-graph --others skip"
timeout 300 python bench.py --workload ecog --steps 1 --warmup 1 $B > gpurun_out/r2w_plain_ecog.json 2> gpurun_out/r2w_plain_ecog.err; echo "ecog rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2w_launches_ecog.csv python bench.py --workload ecog --steps 1 --warmup 1 $B > gpurun_out/r2w_ncu_launch_ecog.log 2>&1; echo "launchlist rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:"k_latent_fused|k_gram_mma" -s 17 -c 2 -o gpurun_out/r2w_prof_ecog python bench.py --workload ecog --steps 1 --warmup 1 $B > gpurun_out/r2w_ncu_ecog.log 2>&1; echo "ncu ecog rc=$?"
ls -la gpurun_out/ | tail -8
