# round-2 check H: decode active box (parity + timing)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "decode or engine_infer or golden_384" > gpurun_out/t_h.log 2>&1; echo "tests rc=$?"; tail -8 gpurun_out/t_h.log
python - <<'PY'
import sys, json, torch
sys.path.insert(0, '.')
import bench
from sgtapose_b200 import decode
dev = torch.device('cuda', 0)
r = bench.decode_roofline(dev)
print('box : ms', round(r['ms'], 4), 'GB/s', round(r['achieved'], 1), 'frac', round(r['frac'], 4), 'rechecks', r['float64_rechecked_pixels_per_launch'])
decode.full_map(True)
r = bench.decode_roofline(dev)
print('full: ms', round(r['ms'], 4), 'GB/s', round(r['achieved'], 1), 'frac', round(r['frac'], 4))
PY
