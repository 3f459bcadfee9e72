# round-2 check J: fp32 staged bulk-store epilogue + shift-kernel register relayout
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_planes.py tests/test_gpu_ops.py -m gpu -x -q -k "planes or conv or dcn or engine or superpixel or small" > gpurun_out/t_j.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/t_j.log
timeout 120 python tools/conv_bench.py 2 2>&1 | tail -5
timeout 120 python tools/stem_bench.py 2 2>&1 | tail -3
for mode in fp32 bf16; do
timeout -k 5 200 python bench.py --mode $mode --no-cpu-baseline --no-extras 2>gpurun_out/bench_j.err | tee gpurun_out/bench_j_${mode}.json | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$mode', 'frames/s', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'dcn ms', round(d['roofline']['ms_per_step'],3), d.get('parity_checked',{}).get('ok'))
for k,v in list(d.get('kernel_families',{}).get('per_step',{}).items())[:9]: print('   ', k, v)"
tail -3 gpurun_out/bench_j.err
done
