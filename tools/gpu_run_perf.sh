# quick perf check: bench line (no extras) + per-kernel launch list of one un-graphed step
mkdir -p gpurun_out
MODE=${1:-fp32}
timeout -k 5 200 python bench.py --mode $MODE --no-cpu-baseline --no-extras 2>gpurun_out/bench_quick.err | tee gpurun_out/bench_quick.json | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('frames/s', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'dcn ms', round(d['roofline']['ms_per_step'],3), 'frac', round(d['roofline']['frac'],4))"
timeout -k 5 200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --mode $MODE --profile-pass --no-cpu-baseline > gpurun_out/launches.log 2>&1
python tools/launch_summary.py gpurun_out/launches.csv 200 > gpurun_out/launch_summary.txt 2>&1; head -12 gpurun_out/launch_summary.txt
