mkdir -p gpurun_out
timeout 200 ncu --set full --clock-control none --import-source on -k regex:decode_peaks_f32 -c 1 -f -o /tmp/dec python tools/decode_bench.py --ncu > gpurun_out/dec_ncu.log 2>&1
ncu -i /tmp/dec.ncu-rep --page raw --csv > gpurun_out/dec_raw.csv 2>/dev/null
ncu -i /tmp/dec.ncu-rep --page source --csv --launch-skip 0 --launch-count 1 2>/dev/null | gzip -9 > gpurun_out/dec_src_f32.csv.gz
ls -la gpurun_out/dec_*
