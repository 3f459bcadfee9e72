# round-2 check K: lanes-as-rows level-0 attention (parity + timing vs the lanes-as-keys kernel)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_planes.py -m gpu -x -q -k "attention or encoder or engine_golden or saturates" > gpurun_out/t_k.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/t_k.log
for v in rows keys; do
if [ $v = keys ]; then export SGTA_ATTN_KEYS=1; else unset SGTA_ATTN_KEYS; fi
timeout -k 5 200 python bench.py --no-cpu-baseline --no-extras 2>gpurun_out/bench_k.err | tee gpurun_out/bench_k_$v.json | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$v', 'frames/s', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d.get('parity_checked',{}).get('ok'), 'attn', d['kernel_families']['per_step']['attn_forward'], 'convs frac', round(d['roofline_convs']['frac'],4))"
tail -3 gpurun_out/bench_k.err
done
