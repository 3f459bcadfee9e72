"""Time single planes convolutions (CUDA events) with the debug toggles of sgta_debug_flags."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sgtapose_b200 import planes as P, _lib
DEV = "cuda"
def bench(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
cfgs = [(64, 64, 64, 96, 3), (64, 128, 128, 48, 3), (64, 256, 256, 24, 3), (32, 64, 768, 96, 3), (64, 128, 64, 96, 1)]
ns_list = [int(a) for a in sys.argv[1:]] or [2, 1]
NTF = int(os.environ.get("NT64", "0")) * 64 | int(os.environ.get("DBG", "0"))
for ns in ns_list:
    for B, Ci, Co, H, k in cfgs:
        _lib.load().sgta_debug_flags(NTF)
        xb = P.PlaneBuf(B, Ci, H, H, ns, DEV); xb.t.random_(-3000, 3000)
        yb = P.PlaneBuf(B, Co, H, H, ns, DEV)
        w = torch.randn(Co, Ci, k, k, device=DEV) * 0.05
        spec = P.ConvSpec(P.weight_matrix(w), torch.ones(Co, device=DEV), torch.zeros(Co, device=DEV), Ci, k, 1, ns, P.ACT_RELU)
        fl = 2.0 * B * H * H * Co * Ci * k * k
        out = []
        for flags in (0, 8192, 4, 7):
            _lib.load().sgta_debug_flags(flags | NTF)
            us = bench(lambda: P.conv(spec, xb.full, yb.full))
            out.append("f%d %7.1fus %6.1fTF" % (flags, us, fl / us / 1e6))
        _lib.load().sgta_debug_flags(0)
        print("ns%d B%d %d->%d @%d k%d | " % (ns, B, Ci, Co, H, k) + " | ".join(out), flush=True)
