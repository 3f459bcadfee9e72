# Round-end style validation on the GPU box: every GPU parity test, smoke(), the bench line, the reference arm,
# and the ncu launch list of one un-graphed step.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t_all.log 2>&1; echo "gpu tests rc=$?"; tail -4 gpurun_out/t_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_fp32.json 2> gpurun_out/bench_fp32.err; echo "bench rc=$?"; cut -c1-700 gpurun_out/bench_fp32.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref rc=$?"; cat gpurun_out/bench_reference.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --profile-pass --no-cpu-baseline > gpurun_out/launches.log 2>&1; echo "ncu rc=$?"
python tools/launch_summary.py gpurun_out/launches.csv 200 > gpurun_out/launch_summary.txt 2>&1; head -12 gpurun_out/launch_summary.txt
