mkdir -p gpurun_out
bash tools/ncu_capture.sh attn "attn_fwd_rows" 1 1 > /dev/null 2>&1
ls -la gpurun_out | grep -i "attn_"
