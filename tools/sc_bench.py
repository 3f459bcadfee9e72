"""Time the small-channel (SC) convolutions of the DLA entry with the debug toggles."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sgtapose_b200 import planes as P, _lib
DEV = "cuda"
def bench(fn, n=10):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
B, S = 64, 384
for ns in (2, 1):
    in4 = P.PlaneBuf(B, 4, S, S, ns, DEV, border=3); in4.t.random_(-3000, 3000)
    f0 = P.PlaneBuf(B, 16, S, S, ns, DEV); l0 = P.PlaneBuf(B, 16, S, S, ns, DEV)
    wi, wh = torch.randn(16, 3, 7, 7, device=DEV) * 0.1, torch.randn(16, 1, 7, 7, device=DEV) * 0.1
    stem = P.ScConvSpec([(wi, 0), (wh, 3)], torch.ones(32, device=DEV), torch.zeros(32, device=DEV), 4, 7, 1, 3, 3, S, ns)
    w0 = torch.randn(16, 16, 3, 3, device=DEV) * 0.1
    lv0 = P.ScConvSpec([(w0, 0)], torch.ones(16, device=DEV), torch.zeros(16, device=DEV), 16, 3, 1, 1, 1, S, ns, P.ACT_RELU)
    for name, fn in (("stem", lambda: P.conv_sc(stem, in4.full, f0.full, P.EPI_STEM)),
                     ("level0", lambda: P.conv_sc(lv0, f0.full, l0.full, P.EPI_SC))):
        out = []
        for flags in (0, 1, 4, 5, 16, 21):
            _lib.load().sgta_debug_flags(flags)
            out.append("f%d %7.1fus" % (flags, bench(fn)))
        _lib.load().sgta_debug_flags(0)
        print("ns%d %-6s | " % (ns, name) + " | ".join(out), flush=True)
