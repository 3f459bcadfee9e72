# round-2 check U: 2-D tiles (8 x 16 pixels) for the fp32 __ldg DCN gather (debug flag 524288 = strips of 128 rows)
mkdir -p gpurun_out
timeout -k 10 400 python -m pytest tests/test_gpu_planes.py tests/test_gpu_ops.py -m gpu -x -q --timeout 100 -k "dcn or engine_golden" > gpurun_out/t_u.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/t_u.log
timeout 120 python tools/wide_bench.py 0 524288 2>&1 | tail -5
for dbg in 0 524288; do
timeout -k 5 150 python bench.py --dbg $dbg --no-cpu-baseline --no-extras 2>gpurun_out/bench_u.err | tee gpurun_out/bench_u_$dbg.json | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('dbg $dbg', 'frames/s', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d.get('parity_checked',{}).get('ok'), 'dcn', d['roofline']['ms_per_step'], d['roofline']['frac'])"
tail -3 gpurun_out/bench_u.err
done
cat gpurun_out/engine_384_seed317_err.json | tr -d '\n '; echo
