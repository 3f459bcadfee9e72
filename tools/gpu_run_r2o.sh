# round-2 check O: f = 2 up-sampler (four outputs per thread), wide K step on long-K 128-wide tiles
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_planes.py tests/test_gpu_ops.py -m gpu -x -q -k "planes or conv or engine_golden or superpixel or dcn" > gpurun_out/t_o.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/t_o.log
for dbg in 0 65536; do
timeout -k 5 200 python bench.py --dbg $dbg --no-cpu-baseline --no-extras 2>gpurun_out/bench_o.err | tee gpurun_out/bench_o_$dbg.json | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('dbg $dbg', 'frames/s', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d.get('parity_checked',{}).get('ok'), 'convs frac', round(d['roofline_convs']['frac'],4), 'bf16', d.get('value_bf16'))
for k,v in list(d.get('kernel_families',{}).get('per_step',{}).items())[:12]: print('   ', k, v)"
tail -3 gpurun_out/bench_o.err
done
