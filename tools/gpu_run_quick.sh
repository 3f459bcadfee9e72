mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_planes.py -x -q -k "upsample" > gpurun_out/t_ups.log 2>&1; echo "ups tests rc=$?"; tail -2 gpurun_out/t_ups.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -k regex:upsample -c 8 --csv --log-file gpurun_out/ups_launches.csv python bench.py --profile-pass --no-cpu-baseline > gpurun_out/launches.log 2>&1; grep "upsample" gpurun_out/ups_launches.csv | awk -F'","' '{print $NF}' | tr '\n' ' '
