"""Time the full-resolution entry layers alone (bench shape: 64 images of 384 x 384) with the debug toggles of
sgta_debug_flags: 1 = producers skip the A loads, 2 = skip the B copies, 4 = skip the epilogue, 16 = skip the MMAs."""
import os, sys
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
for ns in [int(a) for a in sys.argv[1:]] or [2, 1]:
    in4 = P.PlaneBuf(B, 4, S, S, ns, DEV, border=3); in4.t.random_(-3000, 3000)
    f0sp = P.PlaneBuf(B, 64, S, S // 4, ns, DEV)
    l0 = P.PlaneBuf(B, 16, S, S, ns, DEV)
    l1 = P.PlaneBuf(B, 32, S // 2, S // 2, ns, DEV)
    g = lambda *s: torch.randn(*s, device=DEV) * 0.1
    one = lambda n: torch.ones(n, device=DEV)
    stem = P.StemSuperSpec(g(16, 3, 7, 7), g(16, 1, 7, 7), one(32), one(32) * 0.1, S, ns)
    l0s = P.ConvSpec(P.weight_matrix(P.superpixel_weight(g(16, 16, 3, 3), 4, 4, 1)), one(64), one(64) * 0.1, 64, 3, 1, ns, P.ACT_RELU)
    l1s = P.ScConvSpec([(g(32, 16, 3, 3), 0)], one(32), one(32) * 0.1, 16, 3, 2, 1, 1, S, ns, P.ACT_RELU)
    runs = [("stem_sp", lambda: P.conv_stem_sp(stem, in4.full, f0sp.full)),
            ("level0_sp", lambda: P.conv(l0s, f0sp.full, y=l0.full, epi=P.EPI_SP2SC)),
            ("level1", lambda: P.conv_sc(l1s, l0.full, l1.full, P.EPI_SC))]
    for name, fn in runs:
        out = []
        for flags in (0, 1, 2, 4, 16, 1 | 2, 1 | 2 | 4, 1 | 2 | 16, 4 | 16):
            _lib.load().sgta_debug_flags(flags)
            out.append("f%-2d %6.0f" % (flags, bench(fn)))
        _lib.load().sgta_debug_flags(0)
        print("ns%d %-10s us: " % (ns, name) + " | ".join(out), flush=True)
