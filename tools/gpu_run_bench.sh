mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/bench_fp32.json 2> gpurun_out/bench_fp32.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench_fp32.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --profile-pass --no-cpu-baseline > gpurun_out/launches.log 2>&1; echo "ncu rc=$?"
python tools/launch_summary.py gpurun_out/launches.csv 200 > gpurun_out/launch_summary.txt 2>&1; head -30 gpurun_out/launch_summary.txt
