"""Generate tests/golden/*.npz from the UNMODIFIED reference imported from /root/reference.

Run in the build container only:  python -m oracle.make_golden
The reference cannot travel to the GPU box, so its outputs on seeded inputs are committed
here as small fixtures; inputs and weights are regenerated from seeds at test time
(tests/_cases.py, sgtapose_b200/synth.py).  DCN inside the reference modules is the
torchvision stand-in (oracle/ref_import.py::TorchvisionDCN), see oracle/__init__.py.

`.cuda()` calls inside the reference decode (utils.py:114,155,165,270-275) are neutralised by
patching torch.Tensor.cuda to the identity for the duration of the run: no GPU exists here.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_import as R           # noqa: E402
from sgtapose_b200 import synth              # noqa: E402
from tests import _cases as C                # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_grad_enabled(False)
    torch.Tensor.cuda = lambda self, *a, **k: self
    ns = R.load_reference()
    opt = R.default_opt()

    # ---- 1. whole model, S=128, B=2, synthetic weights -------------------------------------
    model = R.build_reference_model(ns, opt)
    sd = synth.synthetic_state_dict(model.state_dict(), seed=C.GOLDEN_SEED)
    model.load_state_dict(sd)
    ins = synth.synthetic_inputs(2, 128, seed=C.GOLDEN_SEED, frame=1)
    cap = {}
    hooks = []
    for i in range(3):
        hooks.append(model.transformer[i].register_forward_hook(
            lambda m, a, o, i=i: cap.__setitem__("tr%d_out" % i, o.numpy().copy())))
    for i in range(6):
        hooks.append(model.cat_layer[i].register_forward_hook(
            lambda m, a, o, i=i: cap.__setitem__("cat%d_rows" % i, o.numpy().copy())))
    hooks.append(model.hm.register_forward_hook(
        lambda m, a, o: cap.__setitem__("feat", a[0].numpy().copy())))
    out = model(*ins)[0]
    for h in hooks:
        h.remove()
    np.savez_compressed(os.path.join(OUT, "model_S128.npz"),
                        hm=out["hm"].numpy(), reg=out["reg"].numpy(), tracking=out["tracking"].numpy(),
                        **cap)
    print("model_S128:", {k: v.shape for k, v in cap.items()})

    # ---- 2. DeformConv blocks ---------------------------------------------------------------
    blobs = {}
    for name, B, Cin, Cout, H, W in C.DEFORMCONV_CASES:
        blk = ns.dla.DeformConv(Cin, Cout).eval()
        blk.load_state_dict(C.deformconv_params(name, Cin, Cout))
        x = C.deformconv_input(name, B, Cin, H, W)
        blobs[name + "_dcn"] = blk.conv(x).numpy()
        blobs[name + "_block"] = blk(x).numpy()
    np.savez_compressed(os.path.join(OUT, "deformconv.npz"), **blobs)
    print("deformconv:", {k: v.shape for k, v in blobs.items()})

    # ---- 3. token index arithmetic (dla.py:898-968) -----------------------------------------
    pm = torch.from_numpy(C.prior_maps_for_index_cases())
    idx = {}
    pre_idx, _ = ns.dla.get_topk_index(pm, pm, 1)
    idx["topk_xy"] = pre_idx.numpy()
    sizes = [384, 192, 96, 48, 24, 12]
    scales = [4, 2, 1, 1 / 2, 1 / 4, 1 / 8]
    kernels = [12, 6, 3, 1, 1, 1]
    for lvl in range(6):
        feats = torch.zeros(pm.shape[0], 2, sizes[lvl], sizes[lvl])
        _, _, fid = ns.dla.get_topk_features_scale(feats, pre_idx, scale_num=scales[lvl], kernel=kernels[lvl])
        idx["fid_l%d" % lvl] = fid.numpy()
    np.savez_compressed(os.path.join(OUT, "token_index.npz"), **idx)
    print("token_index: level-3 id of (47,47) =", idx["fid_l3"][0, 0])

    # ---- 4. decode ----------------------------------------------------------------------------
    hms = C.decode_heatmaps()
    reg, trk = C.decode_reg_tracking(hms.shape[0])
    keys = ["scores", "clses", "xs", "ys", "cts", "cts_wreg", "regs", "tracking"]
    acc = {k: [] for k in keys}
    for n in range(hms.shape[0]):
        o = {"hm": torch.from_numpy(hms[n:n + 1]), "reg": torch.from_numpy(reg[n:n + 1]),
             "tracking": torch.from_numpy(trk[n:n + 1])}
        d = ns.decode.dream_generic_decode(o, K=7, opt=opt)
        for k in keys:
            acc[k].append(d[k].numpy())
    dec = {k: np.concatenate(v, 0) for k, v in acc.items()}
    # alternate decode: _nms + _topk, and SoftArgmaxPavlo, on the blob samples only (noise maps tie)
    heat = torch.from_numpy(hms[:8])
    nm = ns.utils._nms(heat)
    s, i, c, y, x = ns.utils._topk(nm, K=7)
    dec.update(nms=nm.numpy(), topk_scores=s.numpy(), topk_inds=i.numpy(), topk_clses=c.numpy())
    sa = ns.utils.SoftArgmaxPavlo(7)
    dec["softargmax"] = sa(heat).numpy()
    np.savez_compressed(os.path.join(OUT, "decode.npz"), **dec)
    print("decode:", {k: v.shape for k, v in dec.items()})
    print("scores sample", dec["scores"][6], dec["xs"][6], dec["ys"][6])


if __name__ == "__main__":
    main()
