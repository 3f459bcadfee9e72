mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py tests/test_gpu_preprocess.py -x -q -k "decode or pre_process or warp" > gpurun_out/t_decode.log 2>&1; echo "decode+preprocess tests rc=$?"
tail -15 gpurun_out/t_decode.log
timeout 200 python tools/decode_bench.py > gpurun_out/decode_bench.json 2> gpurun_out/decode_bench.err; cat gpurun_out/decode_bench.json
timeout 200 ncu --set full --clock-control none --import-source on -k regex:decode_peaks_f32 -c 1 -f -o /tmp/dec python tools/decode_bench.py --ncu > gpurun_out/dec_ncu.log 2>&1
ncu -i /tmp/dec.ncu-rep --page raw --csv > gpurun_out/dec_raw.csv 2>/dev/null
ncu -i /tmp/dec.ncu-rep --page source --csv --launch-skip 0 --launch-count 1 2>/dev/null | gzip -9 > gpurun_out/dec_src_f32.csv.gz
