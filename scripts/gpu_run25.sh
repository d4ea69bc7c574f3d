mkdir -p gpurun_out
(SNB_DEBUG_RAW=1 timeout 300 python scripts/gpu_raw.py 512 text; SNB_DEBUG_RAW=1 timeout 300 python scripts/gpu_raw.py 512 mixed) 2>&1 | grep -v "^$" | tail -40 | tee gpurun_out/raw_probe.log
