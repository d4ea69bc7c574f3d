set -x
mkdir -p gpurun_out
for lib in ab_noskip.so ab_skip1.so ab_skip2.so ab_noskip.so; do
SNB_LIB=$PWD/nim-snappy_b200/$lib timeout 600 python bench.py --steps 5 --warmup 3 --skip-cpu --skip-configs > gpurun_out/bench_$lib.json 2> gpurun_out/bench_$lib.err
python - <<P
import json
d=json.loads([l for l in open('gpurun_out/bench_$lib.json') if l.startswith('{')][-1])
print('$lib', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'unc', d['uncompress']['value'], d['uncompress']['with_frame_table']['value'], d['config']['verified_vs_oracle'])
P
done
