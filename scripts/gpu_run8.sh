# round-1h: tests, bench, launch list and full ncu captures of the two hot kernels (compress also with warm caches)
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 2 --warmup 1 --skip-cpu --e2e-steps 1 > gpurun_out/bench_under_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 200 --csv --log-file gpurun_out/launches_warm.csv \
  python bench.py --steps 2 --warmup 1 --skip-cpu --e2e-steps 1 > gpurun_out/bench_under_ncu2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:compress_stream_kernel -s 1 -c 1 -o gpurun_out/prof_compress -f \
  python bench.py --steps 1 --warmup 1 --skip-cpu --e2e-steps 1 > gpurun_out/ncu_compress.log 2>&1
timeout 900 ncu --set full --clock-control none --cache-control none --import-source on -k regex:compress_stream_kernel -s 1 -c 1 -o gpurun_out/prof_compress_warm -f \
  python bench.py --steps 1 --warmup 1 --skip-cpu --e2e-steps 1 > gpurun_out/ncu_compress_warm.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decompress_kernel -s 1 -c 1 -o gpurun_out/prof_decompress -f \
  python bench.py --steps 1 --warmup 1 --skip-cpu --e2e-steps 1 > gpurun_out/ncu_decompress.log 2>&1
ls -la gpurun_out | tail -20
