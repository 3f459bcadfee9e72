# every GPU parity test, with a per-test time limit (a hung kernel must not eat the GPU budget)
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -x -q --timeout 150 > gpurun_out/t_all.log 2>&1; echo "gpu tests rc=$?"; tail -6 gpurun_out/t_all.log
