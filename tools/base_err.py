"""Error of the DLA base per level: eager fp32 (cuDNN) and the engine (fp16 hi/lo tensor cores) vs fp64."""
import sys, os, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
from sgtapose_b200 import config, engine, networks, synth
DEV = "cuda"; S = 128; B = 2
def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()
def rms(a, b):
    a, b = a.double(), b.double()
    return ((a - b).pow(2).mean().sqrt() / b.pow(2).mean().sqrt()).item()
opt = config.default_opt()
m = networks.create_model(config.ARCH, dict(config.HEADS), dict(config.HEAD_CONV), opt).eval()
sd = synth.synthetic_state_dict(m.state_dict(), seed=317)
m.load_state_dict(sd); m = m.to(DEV)
ins = [t.to(DEV) for t in synth.synthetic_inputs(B, S, seed=317, frame=1)]
x, pre_img, pre_hm, repro_hm = ins[:4]
with torch.no_grad():
    e32 = m.base(pre_img=pre_img, pre_hm=pre_hm)
    b64 = copy.deepcopy(m.base).double()
    t64 = b64(pre_img=pre_img.double(), pre_hm=pre_hm.double())
eng = engine.InferenceEngine(sd, opt, batch=B, size=S, mode="fp32", device=DEV, use_graph=False)
eng(*ins)
for i in range(6):
    e = eng.buf["l%d" % i].to_nchw()[:B]
    print("l%d  eager-vs-fp64 max %.2e rms %.2e | engine-vs-fp64 max %.2e rms %.2e" % (
        i, rel(e32[i], t64[i]), rms(e32[i], t64[i]), rel(e, t64[i]), rms(e, t64[i])))
