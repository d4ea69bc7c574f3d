mkdir -p gpurun_out
for r in 1 2; do
for v in 0 1; do
echo "== SNB_TUNE_DEC_TAPER=$v"
SNB_TUNE_DEC_TAPER=$v timeout 300 python scripts/gpu_e2e.py 1024 6 2>&1 | grep "e2e uncompress"
done; done | tee gpurun_out/ab_taper.log
