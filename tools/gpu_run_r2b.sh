# round-2 check B: gather-kernel register relayout (parity + timing), full default bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_planes.py tests/test_gpu_ops.py -m gpu -x -q -k "planes or conv or dcn or engine_golden or stem or small" > gpurun_out/t_b.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/t_b.log
( time timeout 600 python bench.py > gpurun_out/bench_fp32.json 2> gpurun_out/bench_fp32.err ) 2>&1 | grep real; echo "bench rc=$?"; tail -3 gpurun_out/bench_fp32.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_fp32.json').read().strip().splitlines()[-1])
print('frames/s', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'dcn ms', round(d['roofline']['ms_per_step'],3), 'frac', round(d['roofline']['frac'],4))
print('bf16', d.get('value_bf16'), d.get('roofline_bf16',{}).get('ms_per_step'), d.get('roofline_bf16',{}).get('frac'))
print('parity', d.get('parity_checked'))
print('pipeline', json.dumps(d.get('pipeline'))[:1500])
for k,v in list(d.get('kernel_families',{}).get('per_step',{}).items())[:14]: print(' ', k, v)
print('cpu', d.get('cpu_baseline'))
PY
