# round-2 check X: token MLP at C = 32 with 640-thread CTAs (8 lanes per token)
mkdir -p gpurun_out
timeout -k 10 200 python -m pytest tests -m gpu -x -q --timeout 100 -k "token or encoder or mlp or fusion" > gpurun_out/t_x.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/t_x.log
timeout -k 5 200 python bench.py --no-cpu-baseline --no-extras 2>gpurun_out/bench_x.err | tee gpurun_out/bench_x.json | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('frames/s', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), d.get('parity_checked',{}).get('ok'))
for k,v in d.get('kernel_families',{}).get('per_step',{}).items():
    if 'token' in k: print('   ', k, v)"
tail -3 gpurun_out/bench_x.err
