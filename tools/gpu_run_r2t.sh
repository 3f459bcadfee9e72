# round-2 check T: fused heads (sgta_planes_conv_heads); SGTA_UNFUSED_HEADS=1 keeps the two-stage form
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_planes.py -m gpu -x -q -k "conv_heads" > gpurun_out/t_t0.log 2>&1; echo "heads test rc=$?"; tail -8 gpurun_out/t_t0.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/t_t.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/t_t.log
for v in fused unfused; do
if [ $v = unfused ]; then export SGTA_UNFUSED_HEADS=1; else unset SGTA_UNFUSED_HEADS; fi
timeout -k 5 200 python bench.py --no-cpu-baseline --no-extras 2>gpurun_out/bench_t.err | tee gpurun_out/bench_t_$v.json | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$v', 'frames/s', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d.get('parity_checked',{}).get('ok'), 'launches', d['gpu_launches'])
for k,v in d.get('kernel_families',{}).get('per_step',{}).items():
    if 'conv' in k: print('   ', k, v)"
tail -3 gpurun_out/bench_t.err
done
cat gpurun_out/engine_384_seed317_err.json | tr -d '\n '; echo
