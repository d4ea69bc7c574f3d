set -x
mkdir -p gpurun_out
timeout 300 python scripts/gpu_index_probe.py 1024 > gpurun_out/index_probe.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2e_frame_index_launches.csv python scripts/gpu_index_probe.py 1024 > gpurun_out/index_probe_ncu.log 2>&1
SNB_DEBUG_DEVUNC=1 timeout 300 python scripts/gpu_devunc_probe.py 1024 > gpurun_out/devunc_probe.log 2>&1
cat gpurun_out/index_probe.log; tail -20 gpurun_out/devunc_probe.log
