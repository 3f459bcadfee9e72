# round-2 check Y: f = 2 up-sampler with all loads up front and 2 input rows per CTA
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests -m gpu -x -q --timeout 100 -k "upsample or engine_golden" > gpurun_out/t_y.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/t_y.log
for dbg in 0 65536; do
timeout -k 5 200 python bench.py --dbg $dbg --no-cpu-baseline --no-extras 2>gpurun_out/bench_y.err | tee gpurun_out/bench_y_$dbg.json | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('dbg $dbg frames/s', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), d.get('parity_checked',{}).get('ok'), d['kernel_families']['per_step']['planes_upsample_add'])"
tail -3 gpurun_out/bench_y.err
done
