mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -k regex:token_mlp -c 9 --csv --log-file gpurun_out/tok_launches.csv python bench.py --profile-pass --no-cpu-baseline > gpurun_out/launches.log 2>&1; grep "token_mlp" gpurun_out/tok_launches.csv | awk -F'","' '{print substr($5,6,28), $NF}' | tr -d '"' | tr '\n' '|'
