# round-2 check S: packed (FFMA2, transposed K / V) lanes-as-rows attention
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -k "attn or attention or encoder or engine or fusion" > gpurun_out/t_s.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/t_s.log
timeout -k 5 200 python bench.py --no-cpu-baseline --no-extras 2>gpurun_out/bench_s.err | tee gpurun_out/bench_s.json | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('frames/s', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d.get('parity_checked',{}).get('ok'))
for k,v in list(d.get('kernel_families',{}).get('per_step',{}).items())[:7]: print('   ', k, v)"
tail -3 gpurun_out/bench_s.err
