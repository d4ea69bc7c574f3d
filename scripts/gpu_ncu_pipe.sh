#!/bin/bash
# full ncu capture (with source) of the v6 compress kernel; SNB_NCU_MIB = corpus size (default 256)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:compress_pipe_kernel -s 1 -c 1 -o gpurun_out/${OUT:-prof_pipe} -f \
   python bench.py --steps 1 --warmup 1 --skip-cpu --e2e-steps 1 --bytes $((${SNB_NCU_MIB:-256}<<20)) > gpurun_out/ncu_pipe.log 2>&1
tail -3 gpurun_out/ncu_pipe.log
