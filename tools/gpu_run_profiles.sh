# round-2 profile set: ncu --set full of the DCN launches per mode (raw metrics + source page of the top ones), the fused
# head convolution (source page), launch lists of one un-graphed step per mode.  (ncu -k matches the kernel's BASE name:
# the plain gathers are captured too and filtered out by name afterwards.)  Every step has its own time limit.
mkdir -p gpurun_out
for mode in fp32 bf16; do
  timeout -k 10 420 bash tools/ncu_capture.sh dcn_$mode "dcn_tile_kernel|conv_gather_kernel" 23 2 --mode $mode > /dev/null 2>&1
  tail -2 gpurun_out/dcn_${mode}_ncu.log
  timeout -k 5 200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 600 --csv --log-file gpurun_out/launches_$mode.csv python bench.py --mode $mode --profile-pass --no-cpu-baseline > gpurun_out/launches.log 2>&1
  python tools/launch_summary.py gpurun_out/launches_$mode.csv 150 > gpurun_out/launch_summary_$mode.txt 2>&1
done
SKIP=46 timeout -k 10 300 bash tools/ncu_capture.sh heads "conv_shift_kernel" 1 1 > /dev/null 2>&1
tail -2 gpurun_out/heads_ncu.log
ls -la gpurun_out | grep "dcn_\|launch\|heads_"
