mkdir -p gpurun_out
for lib in libsnb200.so ab_pf128.so ab_pf256.so ab_pf512.so; do
  echo "== $lib"
  SNB_LIB=$PWD/nim-snappy_b200/$lib timeout 300 python scripts/gpu_devunc_probe.py 1024 2>&1 | tail -3
done 2>&1 | tee gpurun_out/ab_dec_pf.log
