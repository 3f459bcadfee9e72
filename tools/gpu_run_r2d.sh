# round-2 check D: K-step skipping in the super-pixel level0 + launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_planes.py tests/test_gpu_ops.py -m gpu -x -q -k "superpixel or conv_shift or engine_golden" > gpurun_out/t_d.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/t_d.log
for mode in fp32 bf16; do
timeout -k 5 200 python bench.py --mode $mode --no-cpu-baseline --no-extras 2>gpurun_out/bench_${mode}_q.err | tee gpurun_out/bench_${mode}_q.json | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$mode', 'frames/s', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'dcn ms', round(d['roofline']['ms_per_step'],3), 'frac', round(d['roofline']['frac'],4), d.get('parity_checked',{}).get('ok'))"
tail -3 gpurun_out/bench_${mode}_q.err
timeout -k 5 200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 600 --csv --log-file gpurun_out/launches_$mode.csv python bench.py --mode $mode --profile-pass --no-cpu-baseline > gpurun_out/launches.log 2>&1
python tools/launch_summary.py gpurun_out/launches_$mode.csv 150 > gpurun_out/launch_summary_$mode.txt 2>&1; head -60 gpurun_out/launch_summary_$mode.txt
done
