#!/bin/bash
# A/B of differently-built libraries with the low-noise probes (min of N launches, separate processes, interleaved):
#   scripts/gpu_ab_probe.sh compress|decode [MiB] [text|mixed] lib1.so lib2.so ...     (paths relative to nim-snappy_b200/)
# Build variants with e.g.  nvcc <flags of __graft_entry__.NVCC_FLAGS> -DSNB_V5_HINT=0 nim-snappy_b200/csrc/snb_api.cu -o nim-snappy_b200/ab_hint0.so
what="$1"; mib="${2:-1024}"; kind="${3:-text}"; shift 3
mkdir -p gpurun_out
for round in 1 2; do
  for lib in "$@"; do
    if [ "$what" = compress ]; then
      SNB_LIB=$PWD/nim-snappy_b200/$lib timeout 300 python scripts/gpu_compress_probe.py "$mib" "$kind" 8 2>&1 | tail -1
    else
      echo "== $lib"; SNB_LIB=$PWD/nim-snappy_b200/$lib timeout 300 python scripts/gpu_devunc_probe.py "$mib" 2>&1 | tail -3
    fi
  done
done | tee gpurun_out/ab_probe.log
