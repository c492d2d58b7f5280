set -x
mkdir -p gpurun_out
for sl in 1 4; do
NMGP_KRON_SLOTS=$sl timeout 300 python profiles/microbench/kron_determinism.py 12288 8 > gpurun_out/r2F_det_slots$sl.log 2>&1; echo "rc=$?"
cat gpurun_out/r2F_det_slots$sl.log | tail -8
done
