mkdir -p gpurun_out
timeout 600 python scripts/gpu_history_probe.py 2>&1 | tail -8 | tee gpurun_out/history_probe.log
