mkdir -p gpurun_out
SNB_TUNE_DEC_G4_MIN=4096 timeout 900 python bench.py --steps 3 --warmup 3 --skip-cpu > gpurun_out/bench_g4.json 2> gpurun_out/bench_g4.err
python - <<P
import json
d=json.loads([l for l in open('gpurun_out/bench_g4.json') if l.startswith('{')][-1])
u=d['uncompress']; print('unc', u['value'], u['roofline']['kernel_ms'])
for k,v in d['configs'].items(): print(k, v.get('uncompress_value'), v.get('uncompress_kernel_ms'), v.get('roundtrip_ok'))
P
