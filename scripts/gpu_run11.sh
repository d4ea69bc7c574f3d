set -x
mkdir -p gpurun_out
for lib in libsnb200.so ab_fidx512.so ab_fidx256.so ab_fidx128.so ab_fidx256h2.so; do
  echo "== $lib"
  SNB_LIB=$PWD/nim-snappy_b200/$lib timeout 300 python scripts/gpu_index_probe.py 1024 2>&1 | grep "hint=none"
done > gpurun_out/ab_fidx.log 2>&1
cat gpurun_out/ab_fidx.log
