# round-2 check P: programmatic dependent launch across the step (debug flag 131072 = launches without the attribute)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/t_p.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/t_p.log
for dbg in 0 131072 0 131072; do
timeout -k 5 200 python bench.py --dbg $dbg --no-cpu-baseline --no-extras 2>gpurun_out/bench_p.err | tee gpurun_out/bench_p_$dbg.json | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('dbg $dbg', 'frames/s', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d.get('parity_checked',{}).get('ok'), 'convs frac', round(d['roofline_convs']['frac'],4), 'bf16', d.get('value_bf16'))"
tail -3 gpurun_out/bench_p.err
done
cat gpurun_out/engine_384_seed0_err.json gpurun_out/engine_384_seed317_err.json | tr -d '\n '; echo
