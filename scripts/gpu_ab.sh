#!/bin/bash
# A/B of compress kernel variants on the GPU box: scripts/gpu_ab.sh "<algo:warps,...>" [MiB] [text|mixed]
python scripts/gpu_tune.py ${2:-1024} ${3:-text} "$1" > gpurun_out/tune.log 2>&1
tail -12 gpurun_out/tune.log
