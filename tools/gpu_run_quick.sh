mkdir -p gpurun_out
timeout 200 python bench.py --mode bf16 --no-extras --no-cpu-baseline > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_bf16.json') if l.startswith('{')][-1])
print(d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['ms_per_step'])
PY
