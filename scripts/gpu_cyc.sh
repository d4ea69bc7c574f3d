#!/bin/bash
mkdir -p gpurun_out
( SNB_LIB=$PWD/nim-snappy_b200/libsnb200_cyc384.so python scripts/gpu_cyc.py 256 "1+0,4+0,6+0,4+4,4+8"
  SNB_LIB=$PWD/nim-snappy_b200/libsnb200_cyc.so python scripts/gpu_cyc.py 256 "4+8,4+18" ) 2>&1 | grep -v "^oracle" | tee gpurun_out/cyc.log
