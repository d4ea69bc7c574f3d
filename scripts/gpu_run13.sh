set -x
mkdir -p gpurun_out
MIB=1024 KIND=text bash scripts/gpu_ab_libs.sh "5:22" ab_noskip.so libsnb200.so
cp gpurun_out/ab_libs.log gpurun_out/ab_skip_text.log
MIB=1024 KIND=mixed bash scripts/gpu_ab_libs.sh "5:22" ab_noskip.so libsnb200.so
for i in 1 2 3; do
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "concurrent or allocator" 2>&1 | grep -E "recovered|passed|failed|Error" 
done > gpurun_out/stress.log 2>&1
cat gpurun_out/stress.log
