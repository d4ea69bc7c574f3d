#!/bin/bash
# A/B of differently-built libraries: scripts/gpu_ab_libs.sh "<algo:warps,...>" lib1.so lib2.so ...   (paths relative to nim-snappy_b200/)
variants="$1"; shift
for lib in "$@"; do
  echo "== $lib"
  SNB_LIB=$PWD/nim-snappy_b200/$lib python scripts/gpu_tune.py ${MIB:-1024} ${KIND:-text} "$variants" 2>&1 | grep -E "algo=|decompress"
done | tee gpurun_out/ab_libs.log
