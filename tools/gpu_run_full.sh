# Round-end style validation on the GPU box: every GPU parity test, smoke(), the bench line, the reference arm.
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -x -q --timeout 150 > gpurun_out/t_all.log 2>&1; echo "gpu tests rc=$?"; tail -4 gpurun_out/t_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
( time timeout 600 python bench.py > gpurun_out/bench_fp32.json 2> gpurun_out/bench_fp32.err ) 2>&1 | grep real; echo "bench rc=$?"; cut -c1-400 gpurun_out/bench_fp32.json; tail -2 gpurun_out/bench_fp32.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/bench_reference.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_fp32.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value','ms_per_step','value_bf16','value_skip_dead_levels','gpu_launches') if k in d})
print('e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'], d['roofline']['ms_per_step'], 'bf16', d.get('roofline_bf16',{}).get('frac'))
print('parity', d.get('parity_checked'))
print('cpu', d.get('cpu_baseline'))
print('decode', d.get('roofline_decode',{}).get('frac'), 'pre', d.get('roofline_preprocess',{}).get('frac'))
PY
