# round-2 check A: token_linear + DCN tile parity, tile kernel on/off timing (fp32, bf16), super-pixel level0 check
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_planes.py -m gpu -x -q -k "token_linear or engine_golden or dcn or encoder" > gpurun_out/t_a.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/t_a.log
q() { python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1', 'frames/s', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'dcn ms', round(d['roofline']['ms_per_step'],3), 'frac', round(d['roofline']['frac'],4))"; }
for mode in fp32 bf16; do for dbg in 0 32; do
timeout -k 5 200 python bench.py --mode $mode --dbg $dbg --no-cpu-baseline --no-extras 2>gpurun_out/bench_${mode}_$dbg.err | tee gpurun_out/bench_${mode}_$dbg.json | q "$mode dbg=$dbg"
done; done
timeout 300 python tools/superpixel_level0_check.py > gpurun_out/superpixel_l0.log 2>&1; echo "sp rc=$?"; tail -12 gpurun_out/superpixel_l0.log
timeout -k 5 200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --dbg 32 --profile-pass --no-cpu-baseline > gpurun_out/launches.log 2>&1
python tools/launch_summary.py gpurun_out/launches.csv 200 > gpurun_out/launch_summary.txt 2>&1; head -30 gpurun_out/launch_summary.txt
