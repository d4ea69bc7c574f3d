#!/bin/bash
# One gpurun call: parity tests, smoke, bench, ncu launch list and full captures.  Outputs -> gpurun_out/.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/nproc.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
cat gpurun_out/bench.json
if [ "${SNB_NCU:-1}" = "1" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 2 --warmup 1 --skip-cpu --e2e-steps 1 > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:compress_stream_kernel -s 1 -c 1 -o gpurun_out/prof_compress -f \
   python bench.py --steps 1 --warmup 1 --skip-cpu --e2e-steps 1 --bytes $((256<<20)) > gpurun_out/ncu_compress.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decompress_kernel -s 1 -c 1 -o gpurun_out/prof_decompress -f \
   python bench.py --steps 1 --warmup 1 --skip-cpu --e2e-steps 1 --bytes $((256<<20)) > gpurun_out/ncu_decompress.log 2>&1
fi
ls -la gpurun_out
