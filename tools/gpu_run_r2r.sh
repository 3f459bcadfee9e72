# round-2 check R: scale / shift staging in the bf16 epilogues as well
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/t_r.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/t_r.log
timeout -k 5 400 python bench.py --no-cpu-baseline 2>gpurun_out/bench_r.err | tee gpurun_out/bench_r.json | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('frames/s', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d.get('parity_checked',{}).get('ok'), 'convs frac', round(d['roofline_convs']['frac'],4), 'bf16', d.get('value_bf16'), d.get('roofline_bf16',{}).get('frac'), 'dcn frac', d['roofline']['frac'])
print({k: d[k] for k in d if k.startswith('roofline_') or k.startswith('value_')})
print(d.get('pipeline'))"
tail -3 gpurun_out/bench_r.err
