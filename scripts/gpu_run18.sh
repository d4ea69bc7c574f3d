mkdir -p gpurun_out
SNB_TUNE_DEC_G4_MIN=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:decompress_g4 -s 1 -c 1 -o gpurun_out/g4_prof -f python scripts/gpu_devunc_probe.py 1024 > gpurun_out/g4_ncu.log 2>&1
tail -3 gpurun_out/g4_ncu.log
