"""Time the planes DCN kernel (CUDA events) with the debug toggles of sgta_debug_flags:
1 = no A gathers, 4 = no epilogue, 16 = no MMAs, 32 = gathers without global loads."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sgtapose_b200 import planes as P, _lib
DEV = "cuda"
def bench(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
cfgs = [(32, 64, 64, 96), (32, 128, 64, 48), (32, 128, 128, 48), (32, 256, 128, 24), (32, 512, 256, 12)]
ns_list = [int(a) for a in sys.argv[1:]] or [2, 1]
for ns in ns_list:
    for B, Ci, Co, H in cfgs:
        xb = P.PlaneBuf(B, Ci, H, H, ns, DEV)
        xb.from_nchw(torch.randn(B, Ci, H, H, device=DEV))
        yb = P.PlaneBuf(B, Co, H, H, ns, DEV)
        w = torch.randn(Co, Ci, 3, 3, device=DEV) * 0.05
        spec = P.ConvSpec(P.weight_matrix(w), torch.ones(Co, device=DEV), torch.zeros(Co, device=DEV), Ci, 3, 1, ns, P.ACT_RELU)
        om = torch.zeros(B * (H + 2) * (H + 2) + 256, 32, device=DEV)
        om[:, :18] = (torch.rand(om.shape[0], 18, device=DEV) - 0.5) * 3.0
        om[:, 18:27] = torch.randn(om.shape[0], 9, device=DEV)
        fl = 2.0 * B * H * H * Co * Ci * 9
        out = []
        for flags in (0, 1, 4, 16):
            _lib.load().sgta_debug_flags(flags)
            us = bench(lambda: P.dcn(xb.full, om, spec, spec.scale, spec.shift, yb.full))
            out.append("f%d %6.1fus" % (flags, us))
        _lib.load().sgta_debug_flags(0)
        print("ns%d B%d %d->%d @%d %5.1fTF | " % (ns, B, Ci, Co, H, fl / float(out[0].split()[1][:-2]) / 1e6) + " | ".join(out), flush=True)
