#!/bin/bash
# One gpurun call: launch list of a bench run + full ncu captures (with source) of the two hot kernels.
set -x
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 2 --warmup 1 --skip-cpu --e2e-steps 1 > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:compress_stream_kernel -s 1 -c 1 -o gpurun_out/prof_compress -f \
   python bench.py --steps 1 --warmup 1 --skip-cpu --e2e-steps 1 --bytes $((${SNB_NCU_MIB:-256}<<20)) > gpurun_out/ncu_compress.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decompress_kernel -s 1 -c 1 -o gpurun_out/prof_decompress -f \
   python bench.py --steps 1 --warmup 1 --skip-cpu --e2e-steps 1 --bytes $((${SNB_NCU_MIB:-256}<<20)) > gpurun_out/ncu_decompress.log 2>&1
ls -la gpurun_out
