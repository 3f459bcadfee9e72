"""GPU check for the experimental super-pixel level0 path (InferenceEngine(superpixel_level0=True)).
NOT part of the pytest suite: the path has not run on a GPU yet.  Run on the B200 box:
    python tools/superpixel_level0_check.py
It (1) round-trips the layout copy, (2) compares the level-0 feature map and the heads of the two engines on the golden
inputs (same fp32 hi/lo arithmetic, different summation order: expect ~1e-6 on level0, < 1e-3 on the heads vs the
reference golden), (3) times both engines at the bench shape."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from sgtapose_b200 import config, engine, networks, planes as P, synth  # noqa: E402

if __name__ == "__main__":
    dev = torch.device("cuda", 0)
    opt = config.default_opt()
    tmpl = networks.create_model(config.ARCH, dict(config.HEADS), dict(config.HEAD_CONV), opt)
    sd = synth.synthetic_state_dict(tmpl.state_dict(), seed=317)

    # (1) layout copy round trip against the NCHW view
    sc = P.PlaneBuf(2, 16, 32, 48, 2, dev)
    sp = P.PlaneBuf(2, 64, 32, 12, 2, dev)
    x = torch.randn(2, 16, 32, 48, device=dev)
    sc.from_nchw(x)
    P.superpixels(sc.full, sp.full, True)
    got = P.from_superpixels(sp.to_nchw(), 4)
    print("sc -> super-pixel view max diff", (got - sc.to_nchw()).abs().max().item())
    sc2 = P.PlaneBuf(2, 16, 32, 48, 2, dev)
    P.superpixels(sc2.full, sp.full, False)
    print("round trip exact:", torch.equal(sc2.to_nchw(), sc.to_nchw()))

    # (2) engines on the golden inputs
    gold = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "model_S128.npz"))
    ins = [t.to(dev) for t in synth.synthetic_inputs(2, 128, seed=317, frame=1)]
    outs = {}
    for flag in (False, True):
        eng = engine.InferenceEngine(sd, opt, batch=2, size=128, mode="fp32", device=dev, superpixel_level0=flag,
                                     use_graph=False)
        out = eng(*ins)[0]
        outs[flag] = ({k: v.clone() for k, v in out.items()}, eng.buf["l0"].to_nchw().clone())
    l0a, l0b = outs[False][1], outs[True][1]
    print("level0 map: max |default - superpixel| / max |default| = %.2e" % ((l0a - l0b).abs().max() / l0a.abs().max()).item())
    for k in ("hm", "reg", "tracking"):
        g = torch.from_numpy(gold[k]).to(dev)
        for flag in (False, True):
            o = outs[flag][0][k]
            print(k, "superpixel" if flag else "default   ", "rel err vs reference golden %.2e" % ((o - g).abs().max() / g.abs().max()).item())

    # (3) timing at the bench shape
    ins = [t.to(dev) for t in synth.synthetic_inputs(32, 384, seed=317, frame=1)]
    for flag in (False, True):
        eng = engine.InferenceEngine(sd, opt, batch=32, size=384, mode="fp32", device=dev, superpixel_level0=flag)
        for _ in range(3):
            eng(*ins)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            eng.forward_static()
        e1.record()
        torch.cuda.synchronize()
        print("superpixel_level0=%s: %.3f ms per step" % (flag, e0.elapsed_time(e1) / 10))
        del eng
