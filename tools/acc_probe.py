"""How accurate is the fp16 hi/lo tensor-core convolution as K grows? (accumulator rounding)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
from sgtapose_b200 import planes as P
DEV = "cuda"
g = torch.Generator().manual_seed(0)
for Ci, Co, H, positive in [(64, 64, 24, False), (256, 128, 12, False), (512, 128, 12, False), (512, 128, 12, True)]:
    x = torch.randn(2, Ci, H, H, generator=g)
    w = torch.randn(Co, Ci, 3, 3, generator=g) * (1.0 / (Ci * 9)) ** 0.5
    if positive:
        x, w = x.abs(), w.abs()
    ref = F.conv2d(x.double(), w.double(), None, 1, 1)
    spec = P.ConvSpec(P.weight_matrix(w.to(DEV)), torch.ones(Co, device=DEV), torch.zeros(Co, device=DEV), Ci, 3, 1, 2)
    xb = P.PlaneBuf(2, Ci, H, H, 2, DEV).from_nchw(x.to(DEV))
    yb = P.PlaneBuf(2, Co, H, H, 2, DEV)
    P.conv(spec, xb.full, yb.full)
    got = yb.to_nchw().cpu().double()
    cud = F.conv2d(x.to(DEV), w.to(DEV), None, 1, 1).cpu().double()
    xq = xb.to_nchw().cpu().double()
    refq = F.conv2d(xq, w.double(), None, 1, 1)
    for name, t in (("umma f16x2", got), ("cudnn fp32", cud)):
        e = (t - ref)
        print("K=%d pos=%d %-11s max|e|/max|ref| %.2e  rms(e)/rms(ref) %.2e  mean(e)/mean|ref| %+.2e" % (
            Ci * 9, positive, name, e.abs().max() / ref.abs().max(), e.pow(2).mean().sqrt() / ref.pow(2).mean().sqrt(),
            e.mean() / ref.abs().mean()))
    print("   input quantisation alone: %.2e" % ((refq - ref).abs().max() / ref.abs().max()))
