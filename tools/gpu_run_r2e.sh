mkdir -p gpurun_out
bash tools/ncu_capture.sh tmlp "token_mlp_kernel" 9 3 > /dev/null 2>&1
bash tools/ncu_capture.sh ups "pl_upsample_add_kernel" 8 1 > /dev/null 2>&1
ls -la gpurun_out | grep -i "tmlp_\|ups_"
