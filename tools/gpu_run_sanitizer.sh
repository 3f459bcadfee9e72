# compute-sanitizer (memcheck, then racecheck on the shared-memory heavy kernels) over the round-2 kernels' parity tests
mkdir -p gpurun_out
K="token_linear or superpixel or post_process or active_box or (dcn_planes and (cfg0 or cfg5 or cfg6)) or conv_shift_vs_torch and (cfg0 or cfg4) or attention_core_vs_torch and 63"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 5 python -m pytest tests/test_gpu_planes.py tests/test_gpu_ops.py tests/test_gpu_detector.py -m gpu -x -q -k "$K" > gpurun_out/san_mem.log 2>&1; echo "memcheck rc=$?"; tail -6 gpurun_out/san_mem.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 5 python -m pytest tests/test_gpu_ops.py tests/test_gpu_detector.py -m gpu -x -q -k "token_linear or active_box or post_process_device" > gpurun_out/san_race.log 2>&1; echo "racecheck rc=$?"; tail -6 gpurun_out/san_race.log
