mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2e_bench_2gpu.json 2> gpurun_out/r2e_bench_2gpu.err; echo "rc $?"
tail -c 600 gpurun_out/r2e_bench_2gpu.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2e_bench_reference_arm_2gpu.json 2> gpurun_out/r2e_ref2.err; echo "rc $?"
