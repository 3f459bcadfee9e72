# round-2 check G: resident weights in the shift kernel (A/B via debug flag 2048)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_planes.py tests/test_gpu_ops.py -m gpu -x -q -k "conv_shift or superpixel or dcn_planes or engine_golden or attention" > gpurun_out/t_g.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/t_g.log
for mode in fp32 bf16; do for dbg in 0 2048; do
timeout -k 5 200 python bench.py --mode $mode --dbg $dbg --no-cpu-baseline --no-extras 2>gpurun_out/bench_g_$dbg.err | tee gpurun_out/bench_g_${mode}_$dbg.json | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$mode dbg $dbg', 'frames/s', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'dcn ms', round(d['roofline']['ms_per_step'],3), d.get('parity_checked',{}).get('ok'))
for k,v in list(d.get('kernel_families',{}).get('per_step',{}).items())[:7]: print('   ', k, v)"
tail -3 gpurun_out/bench_g_$dbg.err
done; done
