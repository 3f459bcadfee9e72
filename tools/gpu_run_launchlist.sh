# launch list of one un-graphed fp32 step (ncu, one metric) + the quick parity subset
mkdir -p gpurun_out
timeout -k 5 200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 600 --csv --log-file gpurun_out/launches_fp32.csv python bench.py --mode fp32 --profile-pass --no-cpu-baseline > gpurun_out/launches.log 2>&1
python tools/launch_summary.py gpurun_out/launches_fp32.csv 150 > gpurun_out/launch_summary_fp32.txt 2>&1; head -4 gpurun_out/launch_summary_fp32.txt
timeout -k 10 300 python -m pytest tests -m gpu -x -q --timeout 100 -k "engine_golden or attention" > gpurun_out/t_ll.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/t_ll.log
