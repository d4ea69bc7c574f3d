# round-2e: tests, bench (new frame-index scan), decode geometry A/B
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 --skip-cpu > gpurun_out/bench_r2e.json 2> gpurun_out/bench_r2e.err; echo "bench exit $?"
for lib in libsnb200.so ab_dec224x4.so ab_dec128x7.so ab_dec160x6.so; do
  echo "== $lib"
  SNB_LIB=$PWD/nim-snappy_b200/$lib timeout 300 python scripts/gpu_devunc_probe.py 1024 2>&1 | tail -3
done > gpurun_out/ab_dec_geom.log 2>&1
cat gpurun_out/ab_dec_geom.log
