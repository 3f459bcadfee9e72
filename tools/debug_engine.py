"""Layer-by-layer comparison of the compiled engine against the eager module tree (GPU)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
from sgtapose_b200 import config, engine, networks, synth

DEV = "cuda"
S = int(sys.argv[1]) if len(sys.argv) > 1 else 128
SEED = int(sys.argv[2]) if len(sys.argv) > 2 else 317
MODES = sys.argv[3].split(",") if len(sys.argv) > 3 else ["fp32", "bf16"]
B = 2
from sgtapose_b200 import dcn_v2
dcn_v2.DCN.tensor_core = False          # the eager tree is the exact-fp32 side of the comparison


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


opt = config.default_opt()
m = networks.create_model(config.ARCH, dict(config.HEADS), dict(config.HEAD_CONV), opt).eval()
sd = synth.synthetic_state_dict(m.state_dict(), seed=SEED)
m.load_state_dict(sd)
m = m.to(DEV)
ins = [t.to(DEV) for t in synth.synthetic_inputs(B, S, seed=SEED, frame=1)]
cap = {}


def hook(name):
    def f(mod, a, o):
        cap.setdefault(name, []).append(o.detach().clone())
    return f


for i in range(6):
    getattr(m.base, "level%d" % i).register_forward_hook(hook("l%d" % i))
m.base.pre_img_layer.register_forward_hook(hook("stem_img"))
m.base.pre_hm_layer.register_forward_hook(hook("stem_hm"))
for name, n in (("dla_up.ida_0", 1), ("dla_up.ida_1", 2), ("dla_up.ida_2", 3), ("ida_up", 2)):
    mod = m
    for part in name.split("."):
        mod = getattr(mod, part)
    for k in range(1, n + 1):
        getattr(mod, "proj_%d" % k).register_forward_hook(hook("%s.%d.p" % (name, k)))
        getattr(mod, "node_%d" % k).register_forward_hook(hook("%s.%d.n" % (name, k)))
        getattr(mod, "node_%d" % k).register_forward_hook(
            lambda mod_, a, o, key="%s.%d.s" % (name, k): cap.setdefault(key, []).append(a[0].detach().clone()))
orig_fuse = m.fuse_level


def fuse(i, *a):
    r = orig_fuse(i, *a)
    cap["fused%d" % i] = [r[0].detach().clone()]
    return r


m.fuse_level = fuse
with torch.no_grad():
    out = m(*ins)[0]

for mode in MODES:
    eng = engine.InferenceEngine(sd, opt, batch=B, size=S, mode=mode, device=DEV, use_graph=False)
    eo = eng(*ins)[0]
    torch.cuda.synchronize()
    print("==== mode", mode)
    nchw = lambda t: t.to_nchw()
    f0 = cap["stem_img"][0] + cap["stem_hm"][0]
    f1 = cap["stem_img"][1] + cap["stem_hm"][1]
    ef0 = nchw(eng.buf["f0"])
    print("stem pre %.3e cur %.3e" % (rel(ef0[:B], f0), rel(ef0[B:], f1)))
    for i in range(6):
        e = nchw(eng.buf["l%d" % i])
        print("l%d pre %.3e" % (i, rel(e[:B], cap["l%d" % i][0])))
    for i in range(6):
        print("fused%d %.3e" % (i, rel(nchw(eng.buf["l%d" % i])[B:], cap["fused%d" % i][0])))
    for key in sorted(k for k in cap if k.endswith((".p", ".s", ".n"))):
        print("%s %.3e" % (key, rel(nchw(eng.buf[key]), cap[key][0])))
    for k in ("hm", "reg", "tracking"):
        print(k, "%.3e" % rel(eo[k], out[k]))
    gname = "model_S128.npz" if (S == 128 and SEED == 317) else "model_S%d_seed%d.npz" % (S, SEED)
    gpath = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", gname)
    if os.path.exists(gpath):
        import numpy as np
        g = np.load(gpath)
        for k in ("hm", "reg", "tracking"):
            gk = torch.from_numpy(g[k]).to(DEV)
            print(k, "engine vs golden %.3e   eager vs golden %.3e" % (rel(eo[k], gk), rel(out[k], gk)))
        if "feat" in g:
            print("feat engine vs golden %.3e" % rel(eng.feat, torch.from_numpy(g["feat"]).to(DEV)))
