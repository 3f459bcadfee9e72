mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_planes.py tests/test_gpu_ops.py -x -q -k "dcn or engine_golden" > gpurun_out/t_dcn.log 2>&1; echo "dcn tests rc=$?"; tail -2 gpurun_out/t_dcn.log
timeout 600 python bench.py --no-cpu-baseline --no-extras > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_quick.json') if l.startswith('{')][-1])
print(d['value'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['ms_per_step'])
PY
tail -3 gpurun_out/bench_quick.err
