mkdir -p gpurun_out
for round in 1 2; do
for lib in ab_noland.so libsnb200.so; do
SNB_LIB=$PWD/nim-snappy_b200/$lib timeout 300 python scripts/gpu_compress_probe.py 1024 text 8 2>&1 | tail -1
done; done | tee gpurun_out/ab_landing_probe.log
for lib in ab_noland.so libsnb200.so; do
SNB_LIB=$PWD/nim-snappy_b200/$lib timeout 300 python scripts/gpu_compress_probe.py 1024 mixed 8 2>&1 | tail -1
done | tee -a gpurun_out/ab_landing_probe.log
