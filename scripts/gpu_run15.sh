mkdir -p gpurun_out
for round in 1 2; do
for lib in ab_noskip.so ab_skip3.so; do
SNB_LIB=$PWD/nim-snappy_b200/$lib timeout 300 python scripts/gpu_compress_probe.py 1024 text 10 2>&1 | tail -1
done; done | tee -a gpurun_out/ab_skip_probe.log
