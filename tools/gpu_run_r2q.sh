# round-2 check Q: stem scale/shift staging, token MLP with two tokens per lane group (debug flag 262144 = one)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -k "planes or conv or engine or superpixel or dcn or token or encoder or mlp or stem" > gpurun_out/t_q.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/t_q.log
for dbg in 0 262144; do
timeout -k 5 200 python bench.py --dbg $dbg --no-cpu-baseline --no-extras 2>gpurun_out/bench_q.err | tee gpurun_out/bench_q_$dbg.json | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('dbg $dbg', 'frames/s', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d.get('parity_checked',{}).get('ok'), 'convs frac', round(d['roofline_convs']['frac'],4))
for k,v in list(d.get('kernel_families',{}).get('per_step',{}).items())[:7]: print('   ', k, v)"
tail -3 gpurun_out/bench_q.err
done
cat gpurun_out/engine_384_seed317_err.json | tr -d '\n '; echo
