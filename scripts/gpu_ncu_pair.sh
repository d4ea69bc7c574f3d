#!/bin/bash
# ncu capture of the v7 pair kernel WITHOUT the SASS-patching sections (half of the warps spin at any time: an
# instrumented spin loop makes the replay take forever); SNB_NCU_MIB = corpus size (default 256)
mkdir -p gpurun_out
SNB_TUNE_ALGO=7 timeout ${NCU_TIMEOUT:-300} ncu --section SpeedOfLight --section WarpStateStats --section SchedulerStats --section ComputeWorkloadAnalysis \
   --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy ${NCU_EXTRA} --clock-control none --import-source on -k regex:compress_pair_kernel -s 1 -c 1 -o gpurun_out/${OUT:-prof_pair} -f \
   python bench.py --steps 1 --warmup 1 --skip-cpu --skip-configs --e2e-steps 1 --bytes $((${SNB_NCU_MIB:-256}<<20)) > gpurun_out/ncu_pair.log 2>&1
tail -3 gpurun_out/ncu_pair.log
