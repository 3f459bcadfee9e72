mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_planes.py tests/test_gpu_ops.py -x -q -k "conv or dcn or engine_golden or engine_bf16 or stem" > gpurun_out/t_dcn.log 2>&1; echo "conv tests rc=$?"; tail -3 gpurun_out/t_dcn.log
timeout 600 python bench.py --no-cpu-baseline --no-extras > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_quick.json') if l.startswith('{')][-1])
print(d['value'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['ms_per_step'])
PY
tail -3 gpurun_out/bench_quick.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -k regex:conv_gather -c 24 --csv --log-file gpurun_out/gather_launches.csv python bench.py --profile-pass --no-cpu-baseline > gpurun_out/launches.log 2>&1; grep "conv_gather" gpurun_out/gather_launches.csv | awk -F'","' '{printf "%s %s | ", substr($5,24,8), $NF}' | tr -d '"'; echo
