# round-2 profile set: ncu --set full of the 16 DCN launches per mode (raw metrics + source page of the top ones),
# launch lists of one un-graphed step per mode
mkdir -p gpurun_out
for mode in fp32 bf16; do
  bash tools/ncu_capture.sh dcn_$mode "dcn_tile_kernel|conv_gather_kernel<0|conv_gather_kernel<\(sgta::PROD\)0" 16 2 --mode $mode > /dev/null 2>&1
  tail -2 gpurun_out/dcn_${mode}_ncu.log
  timeout -k 5 200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 600 --csv --log-file gpurun_out/launches_$mode.csv python bench.py --mode $mode --profile-pass --no-cpu-baseline > gpurun_out/launches.log 2>&1
  python tools/launch_summary.py gpurun_out/launches_$mode.csv 150 > gpurun_out/launch_summary_$mode.txt 2>&1
done
bash tools/ncu_capture.sh shift_fp32 "conv_shift_kernel" 49 2 --mode fp32 > /dev/null 2>&1
ls -la gpurun_out | grep "dcn_\|shift_\|launch"
