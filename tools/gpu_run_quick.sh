mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_ops.py -x -q -k "token_mlp or engine_golden" > gpurun_out/t_tok.log 2>&1; echo "token tests rc=$?"; tail -3 gpurun_out/t_tok.log
timeout 600 python bench.py --no-cpu-baseline --no-extras > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_quick.json') if l.startswith('{')][-1])
print(d['value'], d['e2e']['value'], d['ms_per_step'])
PY
tail -3 gpurun_out/bench_quick.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -k regex:token_mlp -c 9 --csv --log-file gpurun_out/tok_launches.csv python bench.py --profile-pass --no-cpu-baseline > gpurun_out/launches.log 2>&1; grep "token_mlp" gpurun_out/tok_launches.csv | awk -F'","' '{print $5, $(NF-6), $(NF-5), $NF}' | cut -c1-160
