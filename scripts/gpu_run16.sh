mkdir -p gpurun_out
SNB_DEBUG_E2E=1 timeout 300 python scripts/gpu_e2e.py 1024 2 > gpurun_out/e2e_timeline.log 2>&1
tail -60 gpurun_out/e2e_timeline.log
