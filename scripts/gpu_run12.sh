set -x
mkdir -p gpurun_out
timeout 300 python scripts/gpu_index_probe.py 1024 > gpurun_out/index_probe.log 2>&1; cat gpurun_out/index_probe.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 --skip-cpu > gpurun_out/bench_r2e.json 2> gpurun_out/bench_r2e.err; echo "bench exit $?"
