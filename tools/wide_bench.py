"""Time the fp32-mode convolutions and DCN layers with the wide K step on / off (debug flags 16384 / 32768)."""
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
L = _lib.load()
FLAGS = [int(a) for a in sys.argv[1:]] or [0, 16384]
cfgs = [(64, 64, 64, 96, 3), (64, 128, 128, 48, 3), (64, 256, 256, 24, 3), (64, 512, 512, 12, 3), (32, 64, 768, 96, 3),
        (64, 128, 64, 96, 1), (64, 448, 128, 48, 1), (32, 64, 32, 96, 3), (32, 128, 32, 48, 3)]
for B, Ci, Co, H, k in cfgs:
    xb = P.PlaneBuf(B, Ci, H, H, 2, DEV); xb.t.random_(-3000, 3000)
    w = torch.randn(Co, Ci, k, k, device=DEV) * 0.05
    if Co % 64 == 0:
        spec = P.ConvSpec(P.weight_matrix(w), torch.ones(Co, device=DEV), torch.zeros(Co, device=DEV), Ci, k, 1, 2, P.ACT_RELU)
    else:       # the offset / mask convolution: 27 channels padded to 32, fp32 rows out
        spec = P.ConvSpec(P.weight_matrix(w[:27]), torch.ones(27, device=DEV), torch.zeros(27, device=DEV), Ci, k, 1, 2)
    out = []
    if Co % 64 == 0:
        yb = P.PlaneBuf(B, Co, H, H, 2, DEV)
        run = lambda: P.conv(spec, xb.full, yb.full)
    else:
        rows = torch.zeros(B * (H + 2) * (H + 2) + 256, 32, device=DEV)
        run = lambda: P.conv(spec, xb.full, y_f32=rows, ld_f32=32, epi=P.EPI_F32ROWS)
    for f in FLAGS:
        L.sgta_debug_flags(f)
        out.append("f%d %7.1fus" % (f, bench(run)))
    L.sgta_debug_flags(0)
    print("conv B%d %d->%d @%d k%d | " % (B, Ci, Co, H, k) + " | ".join(out), flush=True)
for B, Ci, Co, H in [(32, 64, 64, 96), (32, 128, 64, 48), (32, 128, 128, 48), (32, 256, 128, 24), (32, 512, 256, 12)]:
    xb = P.PlaneBuf(B, Ci, H, H, 2, DEV)
    xb.from_nchw(torch.randn(B, Ci, H, H, device=DEV))
    yb = P.PlaneBuf(B, Co, H, H, 2, DEV)
    w = torch.randn(Co, Ci, 3, 3, device=DEV) * 0.05
    spec = P.ConvSpec(P.weight_matrix(w), torch.ones(Co, device=DEV), torch.zeros(Co, device=DEV), Ci, 3, 1, 2, P.ACT_RELU)
    om = torch.zeros(B * (H + 2) * (H + 2) + 256, 32, device=DEV)
    om[:, :18] = (torch.rand(om.shape[0], 18, device=DEV) - 0.5) * 3.0
    om[:, 18:27] = torch.randn(om.shape[0], 9, device=DEV)
    out = []
    for f in FLAGS:
        L.sgta_debug_flags(f)
        out.append("f%d %7.1fus" % (f, bench(lambda: P.dcn(xb.full, om, spec, spec.scale, spec.shift, yb.full), 10)))
    L.sgta_debug_flags(0)
    print("dcn  B%d %d->%d @%d | " % (B, Ci, Co, H) + " | ".join(out), flush=True)
