#!/bin/bash
# Why 46 ms under ncu vs 23 ms in the bench for compress_stream_kernel?  Same launch, three clocks, with and without
# the persisting-L2 carve-out / access-policy window, outside and under the profiler.
mkdir -p gpurun_out
{
echo "== no profiler"; python scripts/gpu_span_probe.py 1024 3 2>&1 | grep launch
echo "== no profiler, SNB_TUNE_L2_TABLES=0"; SNB_TUNE_L2_TABLES=0 python scripts/gpu_span_probe.py 1024 3 2>&1 | grep launch
for v in default 0; do
  echo "== under ncu (gpu__time_duration only), SNB_TUNE_L2_TABLES=$v"
  if [ $v = 0 ]; then export SNB_TUNE_L2_TABLES=0; else unset SNB_TUNE_L2_TABLES; fi
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:compress_stream_kernel --csv --log-file gpurun_out/bisect_$v.csv python scripts/gpu_span_probe.py 1024 3 2>&1 | grep launch
  grep compress_stream_kernel gpurun_out/bisect_$v.csv | awk -F'","' '{print "   ncu gpu__time_duration: " $NF}'
done
unset SNB_TUNE_L2_TABLES
echo "== under ncu, v6 kernel"; SNB_ALGO=6 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:compress_pipe_kernel --csv --log-file gpurun_out/bisect_v6.csv python scripts/gpu_span_probe.py 1024 3 2>&1 | grep launch
grep compress_pipe_kernel gpurun_out/bisect_v6.csv | awk -F'","' '{print "   ncu gpu__time_duration: " $NF}'
} | tee gpurun_out/ncu_bisect.log
