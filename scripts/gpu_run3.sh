set -x
mkdir -p gpurun_out
SNB_TUNE_VERBOSE=1 SNB_TUNE_L2_TABLES=0 python scripts/gpu_tune.py 1024 text "5:2104,5:1304,5:106" > gpurun_out/l2_off.log 2>&1
SNB_TUNE_VERBOSE=1 SNB_TUNE_L2_TABLES=1 python scripts/gpu_tune.py 1024 text "5:2104,5:1704,5:1304,5:904" > gpurun_out/l2_on.log 2>&1
SNB_TUNE_L2_TABLES=1 SNB_TUNE_L2_RATIO=0.6 python scripts/gpu_tune.py 1024 text "5:2104,5:1704" > gpurun_out/l2_on_r06.log 2>&1
SNB_TUNE_L2_TABLES=1 SNB_TUNE_L2_RATIO=1.0 python scripts/gpu_tune.py 1024 text "5:2104,5:1704" > gpurun_out/l2_on_r10.log 2>&1
SNB_TUNE_L2_TABLES=1 SNB_LIB=$PWD/nim-snappy_b200/lib_t1024.so python scripts/gpu_tune.py 1024 text "5:2904,5:2504" > gpurun_out/l2_on_1024t.log 2>&1
grep -h "snb:\|algo=" gpurun_out/l2_*.log
