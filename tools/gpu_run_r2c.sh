# round-2 check C: super-pixel stem + level0 (parity, engine goldens, timing)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_planes.py tests/test_gpu_ops.py -m gpu -x -q -k "superpixel or small_channel or engine or smoke" > gpurun_out/t_c.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/t_c.log
for mode in fp32 bf16; do
timeout -k 5 200 python bench.py --mode $mode --no-cpu-baseline --no-extras 2>gpurun_out/bench_${mode}_q.err | tee gpurun_out/bench_${mode}_q.json | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$mode', 'frames/s', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'dcn ms', round(d['roofline']['ms_per_step'],3), 'frac', round(d['roofline']['frac'],4), d.get('parity_checked'))
for k,v in list(d.get('kernel_families',{}).get('per_step',{}).items())[:8]: print('   ', k, v)"
tail -3 gpurun_out/bench_${mode}_q.err
done
