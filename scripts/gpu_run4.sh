set -x
mkdir -p gpurun_out
for s in 256 512 1024 2048; do SNB_TUNE_DEC_SLAB=$s python scripts/gpu_e2e.py 1024 3 > gpurun_out/e2e4_$s.log 2>&1; done
grep -h "e2e" gpurun_out/e2e4_*.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
