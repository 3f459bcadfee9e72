# ncu --set full of the first 7 shift-GEMM convolution launches of one un-graphed fp32 step (level0 over super-pixels,
# the three 64->64@96^2 3x3 convs of level2, its two 1x1s, the first 128->128@48^2 3x3 conv): raw metrics + the source
# page of the longest one.  ncu saves / restores the step's 9 GB working set around each of its 39 passes (about 9 s
# per captured launch), so the count stays small.
mkdir -p gpurun_out
timeout -k 10 150 bash tools/ncu_capture.sh convs_fp32 "conv_shift_kernel" 7 1 --mode fp32 > /dev/null 2>&1
tail -2 gpurun_out/convs_fp32_ncu.log
python tools/ncu_table.py gpurun_out/convs_fp32_raw.csv > gpurun_out/convs_fp32_table.txt 2>&1
head -12 gpurun_out/convs_fp32_table.txt
ls -la gpurun_out | grep convs_
