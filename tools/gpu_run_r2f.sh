# round-2 check F: fp32-row gather source of the DCN producers + attention q prefetch
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_planes.py tests/test_gpu_ops.py -m gpu -x -q -k "dcn or attention or encoder or engine_golden" > gpurun_out/t_f.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/t_f.log
for dbg in 0 1024; do
timeout -k 5 200 python bench.py --dbg $dbg --no-cpu-baseline --no-extras 2>gpurun_out/bench_f_$dbg.err | tee gpurun_out/bench_f_$dbg.json | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('dbg $dbg', 'frames/s', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'dcn ms', round(d['roofline']['ms_per_step'],3), 'frac', round(d['roofline']['frac'],4), d.get('parity_checked',{}).get('ok'))
for k,v in list(d.get('kernel_families',{}).get('per_step',{}).items())[:9]: print('   ', k, v)"
tail -3 gpurun_out/bench_f_$dbg.err
done
timeout -k 5 200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 600 --csv --log-file gpurun_out/launches_fp32.csv python bench.py --profile-pass --no-cpu-baseline > gpurun_out/launches.log 2>&1
python tools/launch_summary.py gpurun_out/launches_fp32.csv 250 > gpurun_out/launch_summary_fp32.txt 2>&1; head -48 gpurun_out/launch_summary_fp32.txt
