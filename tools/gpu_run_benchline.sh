# the driver's default bench invocation, timed, with the line's key figures echoed
mkdir -p gpurun_out
( time timeout -k 5 100 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err ) 2>&1 | grep real; tail -2 gpurun_out/bench_final.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_final.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value','ms_per_step','value_bf16','value_skip_dead_levels','gpu_launches') if k in d})
print('e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'], 'bf16', d.get('roofline_bf16',{}).get('frac'))
print('clocks', d['clocks'], 'parity', d.get('parity_checked',{}).get('ok'), 'cpu', d.get('cpu_baseline',{}).get('value'))
print('pipeline', {k: round(v['frames_per_s']) for k, v in d.get('pipeline',{}).items()})
PY
