# round-2 check V: per-warp sampling table in dcn_tile_kernel (bf16 mode), explicit STS in the plain gathers
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests -m gpu -x -q --timeout 150 > gpurun_out/t_v.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/t_v.log
for mode in bf16 fp32; do
timeout -k 5 200 python bench.py --mode $mode --no-cpu-baseline --no-extras 2>gpurun_out/bench_v.err | tee gpurun_out/bench_v_$mode.json | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$mode', 'frames/s', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d.get('parity_checked',{}).get('ok'), 'dcn', d['roofline']['ms_per_step'], d['roofline']['frac'])
for k,v in list(d.get('kernel_families',{}).get('per_step',{}).items())[:4]: print('   ', k, v)"
tail -3 gpurun_out/bench_v.err
done
