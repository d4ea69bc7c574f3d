set -x
mkdir -p gpurun_out
python scripts/gpu_tune.py 1024 text "5:2104,5:1704,5:1304,3:0" > gpurun_out/tune_a.log 2>&1
SNB_LIB=$PWD/nim-snappy_b200/lib_t1024.so python scripts/gpu_tune.py 1024 text "5:2904,5:2504,5:2906" > gpurun_out/tune_b.log 2>&1
python scripts/gpu_tune.py 4096 text "3:0,5:2104" > gpurun_out/tune_c.log 2>&1
for s in 512 1024 2048 4096; do SNB_TUNE_DEC_SLAB=$s python scripts/gpu_e2e.py 1024 3 > gpurun_out/e2e_$s.log 2>&1; done
tail -n 8 gpurun_out/tune_*.log gpurun_out/e2e_*.log
