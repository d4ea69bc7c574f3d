mkdir -p gpurun_out
for lib in lib_base.so libsnb200.so lib_base.so libsnb200.so; do echo "== $lib"; SNB_LIB=$PWD/nim-snappy_b200/$lib python scripts/gpu_e2e.py 1024 4 2>&1 | grep -v oracle; done | tee gpurun_out/e2e_ab.log
