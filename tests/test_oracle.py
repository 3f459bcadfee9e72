"""CPU tests: the oracle against the reference's outputs (golden fixtures), against
scipy / numpy / torchvision themselves, and, in the build container, against the imported
reference."""
import numpy as np
import pytest
import torch

from oracle import dcn as odcn
from oracle import decode as odec
from oracle import model as omodel
from oracle import ref_import
from sgtapose_b200 import networks, synth
from tests import _cases as C


def _model_sd():
    m = networks.create_model("dlapawdl3new_34", dict(ref_import.HEADS), dict(ref_import.HEAD_CONV),
                              ref_import.default_opt())
    return synth.synthetic_state_dict(m.state_dict(), seed=C.GOLDEN_SEED)


def test_dcn_oracle_matches_torchvision():
    from torchvision.ops import deform_conv2d
    torch.manual_seed(1)
    for (B, Ci, Co, H, W, st, pad, dil, dg) in [(2, 6, 5, 9, 11, 1, 1, 1, 1), (1, 8, 4, 12, 7, 2, 1, 1, 2),
                                                (1, 4, 4, 10, 10, 1, 2, 2, 1)]:
        x = torch.randn(B, Ci, H, W); w = torch.randn(Co, Ci, 3, 3); b = torch.randn(Co)
        Ho = (H + 2 * pad - (dil * 2 + 1)) // st + 1
        Wo = (W + 2 * pad - (dil * 2 + 1)) // st + 1
        off = torch.randn(B, 18 * dg, Ho, Wo) * 2.5
        m = torch.rand(B, 9 * dg, Ho, Wo)
        a = odcn.dcn_v2_conv(x, off, m, w, b, st, pad, dil, dg)
        r = deform_conv2d(x, off, w, b, stride=st, padding=pad, dilation=dil, mask=m)
        assert torch.allclose(a, r, atol=1e-5, rtol=1e-5)


def test_deformconv_golden(golden):
    g = golden("deformconv.npz")
    for name, B, Cin, Cout, H, W in C.DEFORMCONV_CASES:
        p = C.deformconv_params(name, Cin, Cout)
        sd = {"blk." + k: v for k, v in p.items()}
        x = C.deformconv_input(name, B, Cin, H, W)
        y = odcn.dcn_forward(x, p["conv.weight"], p["conv.bias"], p["conv.conv_offset_mask.weight"],
                             p["conv.conv_offset_mask.bias"])
        np.testing.assert_allclose(y.numpy(), g[name + "_dcn"], rtol=1e-4, atol=1e-5)
        blk = omodel.deform_conv_block(sd, "blk", x, use_torchvision=False)
        np.testing.assert_allclose(blk.numpy(), g[name + "_block"], rtol=1e-4, atol=1e-5)


def test_gaussian_blur_bit_exact_vs_scipy():
    from scipy.ndimage import gaussian_filter
    rng = np.random.default_rng(0)
    for shape in [(96, 96), (120, 120), (32, 40), (13, 17), (5, 96)]:
        m = rng.random(shape, dtype=np.float32)
        assert np.array_equal(odec.gaussian_blur(m), gaussian_filter(m, sigma=3))


def test_centroid_bit_exact_vs_numpy_average():
    rng = np.random.default_rng(1)
    for _ in range(300):
        m = rng.random((96, 96), dtype=np.float32)
        px, py = (int(v) for v in rng.integers(0, 96, 2))
        wts = np.zeros((5, 5)); iv = np.zeros((5, 5)); jv = np.zeros((5, 5))
        for i in range(-2, 3):
            for j in range(-2, 3):
                if py + i < 0 or py + i >= 96 or px + j < 0 or px + j >= 96:
                    continue
                iv[j + 2, i + 2] = py + i; jv[j + 2, i + 2] = px + j; wts[j + 2, i + 2] = m[py + i, px + j]
        e = (np.average(jv, weights=wts) + 0.4395, np.average(iv, weights=wts) + 0.4395)
        assert odec._centroid(m, px, py) == e


def test_decode_golden(golden):
    g = golden("decode.npz")
    hms = C.decode_heatmaps()
    reg, trk = C.decode_reg_tracking(hms.shape[0])
    d = odec.dream_generic_decode(hms, reg, trk)
    for k in ("xs", "ys", "cts"):
        assert np.array_equal(d[k], g[k]), k
    assert np.array_equal(d["scores"], g["scores"])
    assert np.array_equal(d["clses"], g["clses"])
    np.testing.assert_allclose(d["cts_wreg"], g["cts_wreg"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(d["regs"], g["regs"], rtol=0, atol=1e-6)
    assert np.array_equal(d["tracking"], g["tracking"])
    # alternate decode
    assert np.array_equal(odec.nms(hms[:8]), g["nms"])
    s, i, c, _, _ = odec.topk(odec.nms(hms[:8]), 7)
    # blob maps: NMS leaves many exact ties at 0 only below the K-th value of interest
    assert np.array_equal(s, g["topk_scores"])
    # torch.topk leaves the order of exact ties unspecified: compare indices
    # only where the score is unique within its sample
    gs = g["topk_scores"]
    uniq = np.array([[np.sum(gs[b] == gs[b, k]) == 1 for k in range(gs.shape[1])] for b in range(gs.shape[0])])
    assert uniq.sum() > 30
    assert np.array_equal(i[uniq], g["topk_inds"][uniq]) and np.array_equal(c[uniq], g["topk_clses"][uniq])
    np.testing.assert_allclose(odec.soft_argmax(hms[:8]), g["softargmax"], rtol=1e-4, atol=1e-3)


def test_token_index_golden(golden):
    g = golden("token_index.npz")
    pm = torch.from_numpy(C.prior_maps_for_index_cases())
    xy = omodel.topk_index(pm, 1)
    assert np.array_equal(xy.numpy(), g["topk_xy"])
    sizes = [384, 192, 96, 48, 24, 12]
    kernels = [12, 6, 3, 1, 1, 1]
    for lvl in range(6):
        fid = omodel.window_ids(xy, omodel.SCALE_LIST[lvl], kernels[lvl], sizes[lvl], sizes[lvl])
        assert np.array_equal(fid.numpy(), g["fid_l%d" % lvl]), lvl
    assert g["fid_l3"][0, 0] == 1151          # (47,47) * 1/2 -> row 23, col 47


def test_model_golden(golden):
    g = golden("model_S128.npz")
    sd = _model_sd()
    ins = synth.synthetic_inputs(2, 128, seed=C.GOLDEN_SEED, frame=1)
    out, feats = omodel.forward(sd, *ins, return_feats=True)
    np.testing.assert_allclose(feats["feat"].numpy(), g["feat"], rtol=1e-3, atol=2e-4)
    for k in ("hm", "reg", "tracking"):
        np.testing.assert_allclose(out[0][k].numpy(), g[k], rtol=1e-3, atol=2e-4)
    # per-level attention / MLP rows (levels 0,1 never reach the output, so pin them here)
    pre, cur = ins[4], ins[5]
    for i in range(6):
        B, Cc, H, W = feats["x_cur"][i].shape
        pid = omodel.window_ids(omodel.topk_index(pre, 1), omodel.SCALE_LIST[i], (12, 6, 3, 1, 1, 1)[i], H, W)
        cid = omodel.window_ids(omodel.topk_index(cur, 1), omodel.SCALE_LIST[i], (12, 6, 3, 1, 1, 1)[i], H, W)
        key = omodel.gather_tokens(feats["x_pre"][i], pid)
        qry = omodel.gather_tokens(feats["x_cur"][i], cid)
        o = qry
        if i <= 2:
            for _ in range(3):
                o = omodel.encoder_layer(sd, "transformer.%d.layers.0" % i, o, key)
            np.testing.assert_allclose(o.numpy(), g["tr%d_out" % i], rtol=1e-3, atol=2e-4)
        else:
            o = key
        rows = omodel.cat_mlp(sd, "cat_layer.%d" % i, o, qry)
        np.testing.assert_allclose(rows.numpy(), g["cat%d_rows" % i], rtol=1e-3, atol=2e-4)


@pytest.mark.reference
def test_oracle_model_vs_live_reference():
    ns = ref_import.load_reference()
    ref = ref_import.build_reference_model(ns)
    sd = synth.synthetic_state_dict(ref.state_dict(), seed=5)
    ref.load_state_dict(sd)
    ins = synth.synthetic_inputs(1, 64, seed=5, frame=2)
    with torch.no_grad():
        r = ref(*ins)[0]
    o = omodel.forward(sd, *ins)[0]
    for k in r:
        assert torch.allclose(r[k], o[k], rtol=1e-3, atol=1e-4), k


@pytest.mark.reference
def test_state_dict_keys_match_reference():
    ns = ref_import.load_reference()
    ref = ref_import.build_reference_model(ns)
    mine = networks.create_model("dlapawdl3new_34", dict(ref_import.HEADS), dict(ref_import.HEAD_CONV),
                                 ref_import.default_opt())
    a, b = ref.state_dict(), mine.state_dict()
    assert list(a.keys()) == list(b.keys())
    assert all(a[k].shape == b[k].shape for k in a)


@pytest.mark.reference
def test_reference_dla_binds_product_dcn_through_the_shim():
    """INTEGRATION.md level 1: the UNMODIFIED reference dla.py, importing `DCNv2.dcn_v2` from integration/ by its
    own relative import (dla.py:22), constructs every DeformConv with the product DCN (dla.py:545); the model's
    state-dict keys / shapes equal the ones built with the oracle's stand-in, a synthetic state-dict loads, and --
    no CPU fallback -- a CPU forward raises instead of computing.  (The forward itself needs a GPU and the reference
    tree at once, which no box of this project has; numerics of the operator are the `gpu` tests' job.)"""
    import os
    from sgtapose_b200 import _lib, dcn_v2, synth
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    shim = os.path.join(root, "integration", "sgtapose", "lib", "model", "networks", "DCNv2")
    ns = ref_import.load_reference(dcn_shim_dir=shim)
    assert ns.dla.DCN is dcn_v2.DCN
    model = ref_import.build_reference_model(ns)
    dcns = [m for m in model.modules() if isinstance(m, dcn_v2.DCN)]
    assert len(dcns) == 16
    stand_in = ref_import.build_reference_model(ref_import.load_reference())
    a, b = stand_in.state_dict(), model.state_dict()
    assert list(a.keys()) == list(b.keys()) and all(a[k].shape == b[k].shape for k in a)
    model.load_state_dict(synth.synthetic_state_dict(b, seed=1))
    with pytest.raises(_lib.SgtaError):
        with torch.no_grad():
            model(*synth.synthetic_inputs(1, 64, seed=1, frame=1))


# ------------------------------------------------------------------------------------ prior maps
def _sparse_maps(g, S, name):
    arr = np.zeros(tuple(g["S%d_%s_shape" % (S, name)]), np.float32)
    arr.reshape(-1)[g["S%d_%s_idx" % (S, name)]] = g["S%d_%s_val" % (S, name)]
    return arr


@pytest.mark.parametrize("S", [128, 384])
def test_priors_golden(golden, S):
    """oracle/priors.py and the product's HOST helpers (affine, clip) vs the reference's own
    utilities.py outputs (tests/golden/priors.npz, oracle/make_golden_priors.py): bit-exact."""
    from oracle import priors as OP
    from sgtapose_b200 import priors as PP
    g = golden("priors.npz")
    q = S // 4
    c = np.array([320.0, 180.0], dtype=np.float32)
    t_in, t_out = g["S%d_trans_input" % S], g["S%d_trans_output" % S]
    assert np.array_equal(OP.get_affine_transform(c, 640.0, [S, S]), t_in)
    assert np.array_equal(PP.get_affine_transform(c, 640.0, 0, [q, q]), t_out)
    assert np.array_equal(OP.gaussian_table(), g["gaussian"])
    assert np.array_equal(PP._G, g["gaussian"].astype(np.float32))
    hm, cls = _sparse_maps(g, S, "hm"), _sparse_maps(g, S, "cls")
    kps = g["S%d_kps" % S]
    assert np.array_equal(PP.affine_transform_and_clip(kps, t_in, S, S, 640, 360), g["S%d_centres_in" % S])
    assert np.array_equal(PP.affine_transform_and_clip(kps, t_out, q, q, 640, 360), g["S%d_centres_out" % S])
    for i, kp in enumerate(kps):
        assert np.array_equal(OP.affine_transform_and_clip(kp, t_in, S, S, 640, 360), g["S%d_centres_in" % S][i])
        assert np.array_equal(OP.get_prev_hm_wo_noise(kp, t_in, S, S, 640, 360), hm[i])
        assert np.array_equal(OP.get_prev_hm_wo_noise_cls(kp, 7, t_out, q, q, 640, 360), cls[i])
    assert OP.get_prev_hm_wo_noise(None, t_in, S, S, 640, 360).sum() == float(g["S%d_none_hm_sum" % S])


# ------------------------------------------------------------------------------------ pre-processing
def test_preprocess_oracle_matches_reference_golden(golden):
    """oracle/preprocess.py vs the reference's own pre_process output (oracle/make_golden_preprocess.py)."""
    from oracle import preprocess as opre
    from oracle.make_golden_preprocess import CASES, case_image
    g = golden("preprocess.npz")
    for i, (raw, inp, seed) in enumerate(CASES):
        images, meta, _ = opre.pre_process(case_image(raw, seed), inp[0], inp[1])
        assert np.array_equal(images, g["images_%d" % i]), i                   # float32 bit-exact
        assert np.array_equal(meta["trans_input"], g["trans_input_%d" % i])
        assert np.array_equal(meta["trans_output"], g["trans_output_%d" % i])


def test_warp_affine_oracle_matches_cv2():
    """The integer restatement of cv2.warpAffine (uint8, INTER_LINEAR, BORDER_CONSTANT) vs cv2 itself:
    scaling, rotation + shear, up-sampling with borders, negative source coordinates, 1-channel."""
    import cv2
    from oracle import preprocess as opre
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (90, 160, 3), dtype=np.uint8)
    mats = [np.array([[0.6, 0, 0], [0, 0.6, 21.0]]), np.array([[0.75, 0, 0.3], [0, 0.75, 26.3]]),
            np.array([[0.61, 0.05, -3.3], [-0.04, 0.58, 22.5]]), np.array([[1.7, 0.0, -100.5], [0.0, 1.7, -50.25]]),
            np.array([[-0.9, 0.3, 140.0], [0.2, 1.1, -10.0]])]
    for M in mats:
        for ds in [(96, 96), (131, 57)]:
            assert np.array_equal(opre.warp_affine_u8(img, M, ds), cv2.warpAffine(img, M, ds, flags=cv2.INTER_LINEAR))
    gray = img[:, :, 0].copy()
    assert np.array_equal(opre.warp_affine_u8(gray, mats[2], (64, 64)),
                          cv2.warpAffine(gray, mats[2], (64, 64), flags=cv2.INTER_LINEAR))


def test_preprocess_host_meta_matches_reference_golden(golden):
    """Host side of the product (sgtapose_b200/preprocess.py): same matrices / meta as the reference."""
    import types
    from oracle.make_golden_preprocess import CASES
    from sgtapose_b200 import preprocess as pre
    g = golden("preprocess.npz")
    for i, (raw, inp, _) in enumerate(CASES):
        opt = types.SimpleNamespace(fix_res=True, fix_short=-1, input_h=inp[0], input_w=inp[1], down_ratio=4)
        meta = pre.transform_meta(raw[0], raw[1], opt)
        assert np.array_equal(meta["trans_input"], g["trans_input_%d" % i])
        assert np.array_equal(meta["trans_output"], g["trans_output_%d" % i])
        assert np.array_equal(meta["c"], g["c_%d" % i]) and float(meta["s"]) == float(g["s_%d" % i])
        assert (meta["out_height"], meta["out_width"]) == (inp[0] // 4, inp[1] // 4)
    with pytest.raises(Exception):
        pre.warp_normalize(torch.zeros(1, 8, 8, 3, dtype=torch.uint8), np.eye(2, 3), (8, 8))   # no CPU fallback


def test_preprocess_other_testing_modes_match_reference_golden(golden):
    """`fix_short` and keep-resolution modes of _transform_scale (sgta_detector.py:344-362): host meta of the product
    and the oracle's images (every 4th pixel stored) vs the reference's own pre_process."""
    import types
    from oracle import preprocess as opre
    from oracle.make_golden_preprocess import MODE_CASES, case_image
    from sgtapose_b200 import preprocess as pre
    g = golden("preprocess.npz")
    for j, (raw, over, seed) in enumerate(MODE_CASES):
        opt = types.SimpleNamespace(fix_res=True, fix_short=-1, input_h=384, input_w=384, down_ratio=4, pad=31)
        for k, v in over.items():
            setattr(opt, k, v)
        meta = pre.transform_meta(raw[0], raw[1], opt)
        assert np.array_equal(meta["trans_input"], g["mode_trans_input_%d" % j]), j
        assert np.array_equal(meta["trans_output"], g["mode_trans_output_%d" % j]), j
        sizes = [meta["inp_height"], meta["inp_width"], meta["out_height"], meta["out_width"]]
        assert sizes == g["mode_sizes_%d" % j].tolist(), j
        warped = opre.warp_affine_u8(case_image(raw, seed), meta["trans_input"], (meta["inp_width"], meta["inp_height"]))
        x = opre.normalize(warped, np.full((1, 1, 3), 0.5, np.float32), np.full((1, 1, 3), 0.5, np.float32))
        assert np.array_equal(x.transpose(2, 0, 1)[None][:, :, ::4, ::4], g["mode_images_%d" % j]), j
    with pytest.raises(Exception):
        pre.transform_meta(360, 640, types.SimpleNamespace(fix_res=True, input_h=384, input_w=384), scale=0.5)


# ------------------------------------------------------------------------------------ decode rounding bound
def test_decode_float32_blur_stays_inside_the_band_the_kernel_uses():
    """The production decode kernel blurs in float32 and trusts a comparison only when its margin exceeds
    80 * 2^-24 * (v32 for non-negative maps | max|x| otherwise) per value (csrc/decode.cu).  Emulate that blur in numpy
    float32 (25 sequential multiply-adds per pass; numpy does not fuse them, which can only be less accurate than the
    kernel's FMA chain) and check on representative maps that the distance to the oracle's float64-accumulate /
    float32-store result is inside the band with margin, and that the banded classification never contradicts the
    reference predicate (no false "surely a peak", no false "surely not")."""
    w64 = odec.gaussian_weights()
    wf = w64.astype(np.float32)
    U = np.float32(80.0 * 2.0 ** -24)

    def blur32(m):
        h, w = m.shape
        iy = np.array([[odec._reflect(y - 12 + t, h) for t in range(25)] for y in range(h)])
        ix = np.array([[odec._reflect(x - 12 + t, w) for t in range(25)] for x in range(w)])
        tmp = (m[iy[:, 0], :] * wf[0]).astype(np.float32)
        for t in range(1, 25):
            tmp = (tmp + m[iy[:, t], :] * wf[t]).astype(np.float32)
        out = (tmp[:, ix[:, 0]] * wf[0]).astype(np.float32)
        for t in range(1, 25):
            out = (out + tmp[:, ix[:, t]] * wf[t]).astype(np.float32)
        return out

    rng = np.random.default_rng(0)
    hm, _ = synth.synthetic_heatmaps(2, 7, 96, 96, seed=3, noise=0.02, missing_every=4)
    maps = [hm.numpy()[b, c] for b in range(2) for c in range(7)]
    maps += [rng.random((96, 96), dtype=np.float32), (0.01 + 0.002 * rng.standard_normal((96, 96))).astype(np.float32),
             np.full((96, 96), 0.5, np.float32), rng.standard_normal((40, 17)).astype(np.float32),
             (1e-3 * rng.random((120, 120))).astype(np.float32)]
    worst = 0.0
    for m in maps:
        v, ref = blur32(m), odec.gaussian_blur(m)
        nonneg = m.min() >= 0
        ev = (U * v).astype(np.float32) if nonneg else np.full_like(v, U * np.float32(np.abs(m).max()))
        err = np.abs(v.astype(np.float64) - ref.astype(np.float64))
        assert (err <= 0.5 * ev.astype(np.float64) + 1e-45).all()             # at least 2x inside the band
        worst = max(worst, float((err / np.maximum(ev.astype(np.float64), 1e-300)).max()))
        h, w = m.shape
        pad = np.zeros((h + 2, w + 2), np.float32); pad[1:-1, 1:-1] = v
        pev = np.zeros((h + 2, w + 2), np.float32); pev[1:-1, 1:-1] = ev
        padr = np.zeros((h + 2, w + 2), np.float32); padr[1:-1, 1:-1] = ref
        T = np.float32(0.01)
        no, sure = v < T - ev, v > T + ev
        peak = ref > T
        for sl in ((slice(0, -2), slice(1, -1)), (slice(2, None), slice(1, -1)), (slice(1, -1), slice(0, -2)),
                   (slice(1, -1), slice(2, None))):
            nb, band = pad[sl], ev + (pev[sl] if nonneg else ev)
            inside = np.zeros((h + 2, w + 2), bool); inside[1:-1, 1:-1] = True
            band = np.where(inside[sl], band, ev)                              # neighbours outside the map are exact zeros
            d = v - nb
            no |= d < -band
            sure &= d > band
            peak &= ref >= padr[sl]
        assert not (no & peak).any() and not (sure & ~peak).any()
    assert worst < 0.2                                                         # measured: a few ulps against a bound of 80


def test_oracle_pnp_front_end_vs_reference_golden(golden):
    """oracle/detector.py::is_pnp (cv2.Rodrigues rotation) == the reference's is_pnp outputs (pnp.npz)."""
    from oracle import detector as odet
    from tests import _cases as C
    g = golden("pnp.npz")
    for i, (prev, kps, nxt) in enumerate(C.pnp_cases()):
        good = np.unique(np.where(kps > C.MISSING)[0])
        a, b = odet.is_pnp(prev[good], kps[good], nxt, kps, C.CAMERA_K)
        assert np.array_equal(a, g["prev_%d" % i])
        assert np.allclose(b, g["next_%d" % i], rtol=0, atol=1e-9), i


@pytest.mark.reference
def test_oracle_decode_vs_live_reference_random_maps():
    """The oracle's live decode against the UNMODIFIED reference `dream_generic_decode` (decode.py:184-313 ->
    utils.py:207-284 -> image_proc.py:1032-1143) on freshly seeded maps that are NOT in the golden set: 0-3 blobs per
    channel with random centres (borders included), widths and amplitudes over a noise floor.  `.cuda()` inside the
    reference decode is patched to the identity for the call (no GPU here), as in oracle/make_golden.py."""
    ns = ref_import.load_reference()
    opt = ref_import.default_opt()
    rng = np.random.default_rng(20241)
    N = 12
    yy, xx = np.mgrid[0:96, 0:96].astype(np.float32)
    hms = (rng.random((N, 7, 96, 96), dtype=np.float32) * 0.005).astype(np.float32)
    for n in range(N):
        for c in range(7):
            for _ in range(int(rng.integers(0, 4))):
                cx, cy = rng.uniform(-1, 96, 2)
                s2 = rng.uniform(1.0, 9.0)
                hms[n, c] = np.maximum(hms[n, c], (rng.uniform(0.05, 1.0) *
                                                   np.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / (2 * s2))).astype(np.float32))
    reg = (rng.standard_normal((N, 2, 96, 96)) * 0.3).astype(np.float32)
    trk = (rng.standard_normal((N, 2, 96, 96)) * 2.0).astype(np.float32)
    d = odec.dream_generic_decode(hms, reg, trk)
    saved = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        found = 0
        for n in range(N):
            o = {"hm": torch.from_numpy(hms[n:n + 1]), "reg": torch.from_numpy(reg[n:n + 1]),
                 "tracking": torch.from_numpy(trk[n:n + 1])}
            r = ns.decode.dream_generic_decode(o, K=7, opt=opt)
            for k in ("xs", "ys", "cts", "scores", "clses", "tracking"):
                assert np.array_equal(d[k][n:n + 1], r[k].numpy()), (k, n)
            for k in ("cts_wreg", "regs"):
                np.testing.assert_allclose(d[k][n:n + 1], r[k].numpy(), rtol=0, atol=1e-6)
            found += int((r["scores"].numpy() > 0).sum())
    finally:
        torch.Tensor.cuda = saved
    assert 20 < found < 7 * N                             # both outcomes (a detection / missing or ambiguous) occur


@pytest.mark.reference
def test_oracle_priors_vs_live_reference_random_keypoints():
    """oracle/priors.py and the product's host affine / clip helpers against the UNMODIFIED reference
    sgtapose/utilities.py (:800-853, :889-972, :1045-1098) on 60 freshly seeded keypoint sets per size that are not in
    the golden file (points inside, outside and on the edges of the raw frame): bit-exact."""
    from oracle import priors as OP
    from oracle.make_golden_priors import load_utilities
    from sgtapose_b200 import priors as PP
    U = load_utilities()
    rng = np.random.default_rng(9917)
    c = np.array([320.0, 180.0], dtype=np.float32)
    for S in (384, 128, 480):
        q = S // 4
        t_in, t_out = U.get_affine_transform(c, 640.0, 0, [S, S]), U.get_affine_transform(c, 640.0, 0, [q, q])
        assert np.array_equal(OP.get_affine_transform(c, 640.0, [S, S]), t_in)
        assert np.array_equal(PP.get_affine_transform(c, 640.0, 0, [q, q]), t_out)
        for _ in range(60):
            kp = rng.uniform([-30, -30], [670, 390], size=(7, 2))
            kp[rng.integers(0, 7)] = np.round(kp[rng.integers(0, 7)])             # some exactly on raw pixel centres
            for t, w in ((t_in, S), (t_out, q)):
                want = U.affine_transform_and_clip(kp, t, w, w, 640, 360)
                assert np.array_equal(OP.affine_transform_and_clip(kp, t, w, w, 640, 360), want)
                assert np.array_equal(PP.affine_transform_and_clip(kp[None], t, w, w, 640, 360)[0], want)
            assert np.array_equal(OP.get_prev_hm_wo_noise(kp, t_in, S, S, 640, 360),
                                  U.get_prev_hm_wo_noise(kp, t_in, S, S, 640, 360))
            assert np.array_equal(OP.get_prev_hm_wo_noise_cls(kp, 7, t_out, q, q, 640, 360),
                                  U.get_prev_hm_wo_noise_cls(kp, kp, t_out, q, q, 640, 360))


@pytest.mark.reference
def test_oracle_token_index_vs_live_reference_random_priors():
    """oracle/model.py::topk_index / window_ids / gather_tokens against the UNMODIFIED reference `get_topk_index` and
    `get_topk_features_scale` (dla.py:898-968) on seeded prior maps with unique maxima anywhere in the 96 x 96 map
    (corners and edges included), at all six pyramid levels -- the fp32 index detour of SURVEY.md H4 must come out
    the same for every position, not only for the pinned (47,47) -> 1151 case."""
    ns = ref_import.load_reference()
    rng = np.random.default_rng(4711)
    B = 6
    pm = (rng.random((B, 7, 96, 96), dtype=np.float32) * 0.5).astype(np.float32)
    pos = [(0, 0), (95, 95), (0, 95), (95, 0), (47, 47), (48, 1), (1, 94)]
    for b in range(B):
        for c in range(7):
            x, y = pos[c] if b == 0 else (int(v) for v in rng.integers(0, 96, 2))
            pm[b, c, y, x] = 1.0                                                   # unique maximum (H5: no ties)
    pm = torch.from_numpy(pm)
    ref_xy, _ = ns.dla.get_topk_index(pm, pm, 1)
    xy = omodel.topk_index(pm, 1)
    assert torch.equal(xy, ref_xy)
    sizes, kernels = [384, 192, 96, 48, 24, 12], [12, 6, 3, 1, 1, 1]
    for lvl in range(6):
        feats = torch.from_numpy(rng.standard_normal((B, 3, sizes[lvl], sizes[lvl])).astype(np.float32))
        rows, _, fid = ns.dla.get_topk_features_scale(feats, ref_xy, scale_num=omodel.SCALE_LIST[lvl], kernel=kernels[lvl])
        mine = omodel.window_ids(xy, omodel.SCALE_LIST[lvl], kernels[lvl], sizes[lvl], sizes[lvl])
        assert torch.equal(mine, fid), lvl
        assert torch.equal(omodel.gather_tokens(feats, mine), rows), lvl
