# round-2 profile E: ncu --set full of the level-0 attention kernel and the stem (source-level)
mkdir -p gpurun_out
bash tools/ncu_capture.sh attn "attn_fwd_batched" 1 1 > /dev/null 2>&1
bash tools/ncu_capture.sh stem "conv_gather_kernel<2" 1 1 > /dev/null 2>&1
ls -la gpurun_out | grep -i "attn_\|stem_"
