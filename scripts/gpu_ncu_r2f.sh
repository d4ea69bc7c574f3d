#!/bin/bash
# Round-2f (final, after the decoder instruction diet) evidence in one gpurun call: tests, both bench arms, launch list of a bench run, full ncu captures (with
# source) of the two hot kernels at the bench's size (the capture's kernel time must match roofline.kernel_ms).
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2f_bench_1gpu.json 2> gpurun_out/r2f_bench_1gpu.err; echo "bench exit $?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2f_bench_reference_arm_1gpu.json 2> gpurun_out/r2f_bench_ref.err; echo "ref exit $?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f_launches_bench_steps2.csv \
   python bench.py --steps 2 --warmup 1 --skip-cpu --skip-configs > gpurun_out/r2f_bench_under_ncu.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:compress_stream_kernel -s 1 -c 1 -o gpurun_out/r2f_prof_compress -f \
   python bench.py --steps 1 --warmup 1 --skip-cpu --skip-configs --e2e-steps 1 > gpurun_out/r2f_ncu_compress.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:decompress_kernel -s 1 -c 1 -o gpurun_out/r2f_prof_decompress -f \
   python bench.py --steps 1 --warmup 1 --skip-cpu --skip-configs --e2e-steps 1 > gpurun_out/r2f_ncu_decompress.log 2>&1
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
ls -la gpurun_out | tail -8
