mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_ops.py -x -q -k "attn or attention or engine_golden" > gpurun_out/t_attn.log 2>&1; echo "attn tests rc=$?"; tail -4 gpurun_out/t_attn.log
timeout 600 python bench.py --no-extras --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/bench_quick.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -k regex:attn -c 12 --csv --log-file gpurun_out/attn_launches.csv python bench.py --profile-pass --no-cpu-baseline > gpurun_out/launches.log 2>&1; grep -c attn gpurun_out/attn_launches.csv; grep "attn_fwd_batched" gpurun_out/attn_launches.csv | head -3 | cut -c1-60,200-400
