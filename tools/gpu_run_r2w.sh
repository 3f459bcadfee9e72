# round-2 check W: batch-looping attention kernel for level 1 as well (SGTA_ATTN_POS_MB = pos_embed size threshold in MB)
mkdir -p gpurun_out
for thr in 16 2; do
export SGTA_ATTN_POS_MB=$thr
timeout -k 10 200 python -m pytest tests -m gpu -x -q --timeout 100 -k "attention or encoder or fusion" > gpurun_out/t_w.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/t_w.log
timeout -k 5 200 python bench.py --no-cpu-baseline --no-extras 2>gpurun_out/bench_w.err | tee gpurun_out/bench_w_$thr.json | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('thr $thr', 'frames/s', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), d.get('parity_checked',{}).get('ok'))
for k,v in d.get('kernel_families',{}).get('per_step',{}).items():
    if 'attn' in k: print('   ', k, v)"
tail -3 gpurun_out/bench_w.err
done
