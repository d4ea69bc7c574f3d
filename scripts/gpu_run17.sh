mkdir -p gpurun_out
for v in 1073741824 1; do
echo "== SNB_TUNE_DEC_G4_MIN=$v"
SNB_TUNE_DEC_G4_MIN=$v timeout 300 python scripts/gpu_devunc_probe.py 1024 2>&1 | tail -3
done | tee gpurun_out/ab_g4.log
