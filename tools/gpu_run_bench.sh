mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_detector.py -x -q > gpurun_out/t_detector.log 2>&1; echo "detector tests rc=$?"; tail -15 gpurun_out/t_detector.log
timeout 600 python bench.py > gpurun_out/bench_fp32.json 2> gpurun_out/bench_fp32.err; echo "bench rc=$?"; tail -c 1800 gpurun_out/bench_fp32.json; tail -5 gpurun_out/bench_fp32.err
