mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_full.csv python bench.py --steps 1 --warmup 1 --skip-cpu --e2e-steps 1 --mixed-gib 8 > gpurun_out/bench_full_under_ncu.log 2>&1
tail -c 300 gpurun_out/bench_full_under_ncu.log
