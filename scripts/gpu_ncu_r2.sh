#!/bin/bash
# Round-2 evidence in one gpurun call: launch list of a bench run + full ncu captures (with source) of the two hot
# kernels at the bench's size (1 GiB: the capture's kernel time must match roofline.kernel_ms) + SASS listings.
set -x
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv \
   python bench.py --steps 2 --warmup 1 --skip-cpu --skip-configs > gpurun_out/r2_bench_under_ncu.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:compress_stream_kernel -s 1 -c 1 -o gpurun_out/r2_prof_compress -f \
   python bench.py --steps 1 --warmup 1 --skip-cpu --skip-configs --e2e-steps 1 > gpurun_out/r2_ncu_compress.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:decompress_kernel -s 1 -c 1 -o gpurun_out/r2_prof_decompress -f \
   python bench.py --steps 1 --warmup 1 --skip-cpu --skip-configs --e2e-steps 1 > gpurun_out/r2_ncu_decompress.log 2>&1
ls -la gpurun_out | tail -8
