mkdir -p gpurun_out
for lib in libsnb200.so ab_wdepth2.so ab_ahead2_l1.so ab_ahead2_l2.so libsnb200.so; do
SNB_LIB=$PWD/nim-snappy_b200/$lib timeout 300 python scripts/gpu_compress_probe.py 1024 text 8 2>&1 | tail -1
done | tee gpurun_out/ab_ahead_probe.log
