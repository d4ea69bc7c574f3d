mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2e_bench_8gpu.json 2> gpurun_out/r2e_bench_8gpu.err; echo "rc $?"
tail -c 300 gpurun_out/r2e_bench_8gpu.json
