mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -x -q > gpurun_out/t_all.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/t_all.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
