"""Decode-only leg of bench.py (BASELINE metric: decode HBM GB/s) on its own: prints the roofline object.
`--ncu` runs ONE small launch of each decode kernel for an `ncu --set full` capture."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402

if __name__ == "__main__":
    dev = torch.device("cuda", 0)
    if "--ncu" in sys.argv:
        from sgtapose_b200 import decode, synth
        hm, _ = synth.synthetic_heatmaps(1024, seed=317)
        hm = hm.to(dev)
        decode.peaks_decode(hm)
        decode.peaks_decode(hm, exact64=True)
        torch.cuda.synchronize()
    else:
        print(json.dumps(bench.decode_roofline(dev)))
