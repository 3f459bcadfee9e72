# round-2 check Z: attn_fwd_kernel with the pos_embed loads of a key chunk issued up front
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests -m gpu -x -q --timeout 100 -k "attention or encoder or fusion or engine_golden or training" > gpurun_out/t_z.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/t_z.log
timeout -k 5 200 python bench.py --no-cpu-baseline --no-extras 2>gpurun_out/bench_z.err | tee gpurun_out/bench_z.json | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('frames/s', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), d.get('parity_checked',{}).get('ok'), d['kernel_families']['per_step']['attn_forward'])"
tail -3 gpurun_out/bench_z.err
