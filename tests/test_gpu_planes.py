"""GPU parity tests of the engine's native path: "planes" layouts, TMA shift-GEMM / gather-GEMM
tcgen05 convolutions (incl. DCN) and the memory-bound planes operators, called through the C
ABI, against fp64/fp32 torch and the CPU oracle."""
import pytest
import torch
import torch.nn.functional as F

from oracle import dcn as odcn
from tests import _cases as C

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = {1: 2e-2, 2: 2e-6}          # bf16 planes / fp16 hi+lo planes (fp32-parity mode)


def rel_err(a, b):
    a, b = a.double(), b.double()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


def _buf(x, ns, **kw):
    from sgtapose_b200 import planes as P
    B, Cc, H, W = x.shape
    return P.PlaneBuf(B, Cc, H, W, ns, DEV, **kw).from_nchw(x.to(DEV))


@pytest.mark.parametrize("ns", [1, 2])
@pytest.mark.parametrize("shape", [(2, 64, 9, 11), (1, 192, 5, 7), (3, 16, 6, 6), (2, 32, 8, 4)])
def test_planes_roundtrip(ns, shape):
    x = C.gen(1, *shape) * 3.0
    x[0, 0, 0, 0] = 1e-6
    x[0, 1, 0, 0] = 1234.5678
    back = _buf(x, ns).to_nchw().cpu()
    assert rel_err(back, x) < (4e-3 if ns == 1 else 3e-7)
    if ns == 2:
        assert (back - x).abs().max().item() <= 3e-7 * x.abs().max().item()


def test_planes_fp32_mode_saturates_at_fp16_range():
    """fp32 mode stores hi = fp16(v): |v| beyond 65504 saturates (csrc/planes.cuh clamp_h) instead of turning into
    inf -- stated in include/sgta_b200.h.  Values up to the limit keep ~2^-22 relative accuracy; a convolution whose
    OUTPUT exceeds the range is clamped in its epilogue, one whose inputs sit at 6e4 still sums exactly in fp32."""
    from sgtapose_b200 import planes as P
    x = torch.zeros(1, 64, 4, 4)
    x[0, 0, 0, 0], x[0, 1, 0, 0], x[0, 2, 0, 0], x[0, 3, 0, 0] = 65504.0, 7e4, -1e6, 60000.123
    back = _buf(x, 2).to_nchw().cpu()
    assert back[0, 0, 0, 0] == 65504.0 and back[0, 1, 0, 0] == 65504.0 and back[0, 2, 0, 0] == -65504.0
    assert abs(back[0, 3, 0, 0].item() - 60000.123) < 60000 * 3e-7
    xb = _buf(torch.full((1, 64, 6, 6), 6.0e4), 2)
    w = torch.zeros(64, 64, 1, 1)
    w[0, :2] = 1.0                    # output 1.2e5 -> clamped
    w[1, 0], w[1, 1] = 1.0, -0.5      # output 3.0e4 -> exact
    spec = P.ConvSpec(P.weight_matrix(w.to(DEV)), torch.ones(64, device=DEV), torch.zeros(64, device=DEV), 64, 1, 1, 2)
    yb = P.PlaneBuf(1, 64, 6, 6, 2, DEV)
    P.conv(spec, xb.full, yb.full)
    y = yb.to_nchw().cpu()
    assert torch.all(y[0, 0] == 65504.0) and torch.all(y[0, 1] == 3.0e4) and torch.all(y[0, 2] == 0.0)


SHIFT_CFGS = [  # B, Cin, Cout, H, W, k
    (2, 64, 64, 24, 24, 3),
    (1, 64, 64, 13, 17, 3),
    (3, 128, 128, 12, 12, 3),
    (1, 256, 512, 6, 6, 3),
    (1, 128, 256, 10, 10, 1),
    (2, 448, 128, 12, 12, 1),
    (1, 64, 768, 16, 16, 3),
    (5, 64, 64, 96, 96, 3),
]


@pytest.mark.parametrize("cfg", SHIFT_CFGS)
@pytest.mark.parametrize("ns", [1, 2])
def test_conv_shift_vs_torch(cfg, ns):
    """3x3 / 1x1 stride-1 convolutions (TMA shift-GEMM) with folded scale/shift, residual and
    ReLU; input and output are channel / batch sub-views of wider buffers."""
    from sgtapose_b200 import planes as P
    B, Ci, Co, H, W, k = cfg
    x = C.gen(31, B, Ci, H, W)
    w = C.gen(32, Co, Ci, k, k) * (1.0 / (Ci * k * k)) ** 0.5
    scale = C.gen(33, Co).abs() + 0.5
    shift = C.gen(34, Co) * 0.2
    res = C.gen(35, B, Co, H, W)
    ref = torch.relu(F.conv2d(x.double(), w.double(), None, 1, k // 2) * scale[None, :, None, None].double()
                     + shift[None, :, None, None].double() + res.double())
    spec = P.ConvSpec(P.weight_matrix(w.to(DEV)), scale.to(DEV), shift.to(DEV), Ci, k, 1, ns, P.ACT_RELU)
    xb = P.PlaneBuf(B + 1, Ci + 64, H, W, ns, DEV)           # image 0 and channels [0,64) are decoys
    xb.view(0, 1, 0, 64).from_nchw(C.gen(36, 1, 64, H, W).to(DEV))
    xv = xb.view(1, B, 64, Ci).from_nchw(x.to(DEV))
    rb = _buf(res, ns)
    yb = P.PlaneBuf(B, Co + 64, H, W, ns, DEV)
    P.conv(spec, xv, yb.view(0, B, 64, Co), res=rb.full)
    got = yb.view(0, B, 64, Co).to_nchw().cpu()
    assert float(yb.view(0, B, 0, 64).to_nchw().abs().max()) == 0.0
    assert rel_err(got, ref) < TOL[ns] * (3 if ns == 2 else 1)
    # the zero border must still be zero (the next layer's halo reads it)
    raw = yb.t.view(ns, (Co + 64) // 64, yb.rows, 64)[:, :, yb.guard:yb.guard + B * (H + 2) * (W + 2)]
    raw = raw.view(ns, -1, B, H + 2, W + 2, 64)
    assert int(raw[:, :, :, 0].abs().max()) == 0 and int(raw[:, :, :, :, 0].abs().max()) == 0
    assert int(raw[:, :, :, -1].abs().max()) == 0 and int(raw[:, :, :, :, -1].abs().max()) == 0


@pytest.mark.parametrize("ns", [1, 2])
def test_conv_shift_f32rows_and_nchw(ns):
    from sgtapose_b200 import planes as P
    B, Ci, H, W = 2, 128, 11, 9
    x = C.gen(41, B, Ci, H, W)
    w = C.gen(42, 27, Ci, 3, 3) * 0.03
    b = C.gen(43, 27)
    ref = F.conv2d(x.double(), w.double(), b.double(), 1, 1)
    spec = P.ConvSpec(P.weight_matrix(w.to(DEV)), torch.ones(27, device=DEV), b.to(DEV), Ci, 3, 1, ns)
    xb = _buf(x, ns)
    Pn = B * (H + 2) * (W + 2)
    om = torch.zeros(Pn + 256, 32, device=DEV)
    P.conv(spec, xb.full, y_f32=om, ld_f32=32, epi=P.EPI_F32ROWS)
    got = om[:Pn].view(B, H + 2, W + 2, 32)[:, 1:-1, 1:-1, :27].permute(0, 3, 1, 2).cpu()
    assert rel_err(got, ref) < TOL[ns] * 3
    # 1x1 head: 128 -> 7 with NCHW fp32 output and fused sigmoid
    w2, b2 = C.gen(44, 7, Ci, 1, 1) * 0.1, C.gen(45, 7)
    ref2 = torch.sigmoid(F.conv2d(x.double(), w2.double(), b2.double()))
    spec2 = P.ConvSpec(P.weight_matrix(w2.to(DEV)), torch.ones(7, device=DEV), b2.to(DEV), Ci, 1, 1, ns,
                       P.ACT_SIGMOID, n_valid=7)
    out = torch.zeros(B, 7, H, W, device=DEV)
    P.conv(spec2, xb.full, y_f32=out, epi=P.EPI_NCHW)
    assert rel_err(out.cpu(), ref2) < TOL[ns] * 3


@pytest.mark.parametrize("cfg", [(2, 64, 128, 24, 24), (1, 128, 256, 12, 20), (1, 256, 512, 12, 12)])
@pytest.mark.parametrize("ns", [1, 2])
def test_conv_stride2_vs_torch(cfg, ns):
    from sgtapose_b200 import planes as P
    B, Ci, Co, H, W = cfg
    x = C.gen(51, B, Ci, H, W)
    w = C.gen(52, Co, Ci, 3, 3) * (1.0 / (Ci * 9)) ** 0.5
    scale, shift = C.gen(53, Co).abs() + 0.5, C.gen(54, Co) * 0.2
    ref = torch.relu(F.conv2d(x.double(), w.double(), None, 2, 1) * scale[None, :, None, None].double()
                     + shift[None, :, None, None].double())
    spec = P.ConvSpec(P.weight_matrix(w.to(DEV)), scale.to(DEV), shift.to(DEV), Ci, 3, 2, ns, P.ACT_RELU)
    yb = P.PlaneBuf(B, Co, H // 2, W // 2, ns, DEV)
    P.conv(spec, _buf(x, ns).full, yb.full)
    assert rel_err(yb.to_nchw().cpu(), ref) < TOL[ns] * 3


@pytest.mark.parametrize("ns", [1, 2])
def test_conv_small_channel_layers(ns):
    """The SC-layout convolutions of the DLA entry: dual 7x7 stem, level0, level1 (stride 2),
    level2 entry (32 -> 64 stride 2 and 1x1 projection)."""
    from sgtapose_b200 import planes as P
    B, S = 2, 40
    img, hm = C.gen(61, B, 3, S, S), C.gen(62, B, 1, S, S).abs()
    wi, wh = C.gen(63, 16, 3, 7, 7) * 0.1, C.gen(64, 16, 1, 7, 7) * 0.2
    sc, sh = C.gen(65, 32).abs() + 0.5, C.gen(66, 32) * 0.3
    d = lambda t: t.double()
    ref0 = torch.relu(F.conv2d(d(img), d(wi), None, 1, 3) * d(sc)[None, :16, None, None] + d(sh)[None, :16, None, None]) + \
        torch.relu(F.conv2d(d(hm), d(wh), None, 1, 3) * d(sc)[None, 16:, None, None] + d(sh)[None, 16:, None, None])
    in4 = P.PlaneBuf(B, 4, S, S, ns, DEV, border=3)
    P.pack_stem(img.to(DEV), hm.to(DEV), in4.full, 0)
    stem = P.ScConvSpec([(wi.to(DEV), 0), (wh.to(DEV), 3)], sc.to(DEV), sh.to(DEV), 4, 7, 1, 3, 3, S, ns)
    f0 = P.PlaneBuf(B, 16, S, S, ns, DEV)
    P.conv_sc(stem, in4.full, f0.full, P.EPI_STEM)
    got0 = f0.to_nchw().cpu()
    assert rel_err(got0, ref0) < TOL[ns] * 3
    # level0: 16 -> 16 3x3 s1 (input = our own f0, reference recomputed from it)
    x0 = got0
    w0 = C.gen(67, 16, 16, 3, 3) * 0.1
    s0, t0 = C.gen(68, 16).abs() + 0.5, C.gen(69, 16) * 0.1
    bn = lambda y, s, t: torch.relu(y * d(s)[None, :, None, None] + d(t)[None, :, None, None])
    ref1 = bn(F.conv2d(d(x0), d(w0), None, 1, 1), s0, t0)
    l0spec = P.ScConvSpec([(w0.to(DEV), 0)], s0.to(DEV), t0.to(DEV), 16, 3, 1, 1, 1, S, ns, P.ACT_RELU)
    l0 = P.PlaneBuf(B, 16, S, S, ns, DEV)
    P.conv_sc(l0spec, f0.full, l0.full, P.EPI_SC)
    got1 = l0.to_nchw().cpu()
    assert rel_err(got1, ref1) < TOL[ns] * 3
    # level1: 16 -> 32 3x3 s2
    w1 = C.gen(70, 32, 16, 3, 3) * 0.1
    s1, t1 = C.gen(71, 32).abs() + 0.5, C.gen(72, 32) * 0.1
    ref2 = bn(F.conv2d(d(got1), d(w1), None, 2, 1), s1, t1)
    l1spec = P.ScConvSpec([(w1.to(DEV), 0)], s1.to(DEV), t1.to(DEV), 16, 3, 2, 1, 1, S, ns, P.ACT_RELU)
    l1 = P.PlaneBuf(B, 32, S // 2, S // 2, ns, DEV)
    P.conv_sc(l1spec, l0.full, l1.full, P.EPI_SC)
    got2 = l1.to_nchw().cpu()
    assert rel_err(got2, ref2) < TOL[ns] * 3
    # level2 entry: 32 -> 64 3x3 s2 (PL output) and max-pool + 1x1 projection 32 -> 64 (PL output)
    w2 = C.gen(73, 64, 32, 3, 3) * 0.08
    s2, t2 = C.gen(74, 64).abs() + 0.5, C.gen(75, 64) * 0.1
    ref3 = bn(F.conv2d(d(got2), d(w2), None, 2, 1), s2, t2)
    c2spec = P.ScConvSpec([(w2.to(DEV), 0)], s2.to(DEV), t2.to(DEV), 32, 3, 2, 1, 1, S // 2, ns, P.ACT_RELU)
    y2 = P.PlaneBuf(B, 64, S // 4, S // 4, ns, DEV)
    P.conv_sc(c2spec, l1.full, y2.full, P.EPI_PL)
    assert rel_err(y2.to_nchw().cpu(), ref3) < TOL[ns] * 3
    bot = P.PlaneBuf(B, 32, S // 4, S // 4, ns, DEV)
    P.maxpool2(l1.full, bot.full, 32)
    pooled = bot.to_nchw().cpu()
    assert rel_err(pooled, F.max_pool2d(got2, 2, 2)) < 1e-6
    wp = C.gen(76, 64, 32, 1, 1) * 0.2
    ref4 = F.conv2d(d(pooled), d(wp)) * d(s2)[None, :, None, None] + d(t2)[None, :, None, None]
    pspec = P.ScConvSpec([(wp.to(DEV), 0)], s2.to(DEV), t2.to(DEV), 32, 1, 1, 0, 1, S // 4, ns)
    y3 = P.PlaneBuf(B, 64, S // 4, S // 4, ns, DEV)
    P.conv_sc(pspec, bot.full, y3.full, P.EPI_PL)
    assert rel_err(y3.to_nchw().cpu(), ref4) < TOL[ns] * 3


@pytest.mark.parametrize("ns", [1, 2])
@pytest.mark.parametrize("hw", [(40, 40), (22, 52)])
def test_superpixel_stem_and_level0(ns, hw):
    """Super-pixel forms of the two full-resolution layers (engine default): the dual 7x7 stem as a gather-GEMM over
    groups of 4 output pixels (N = 128, EPI_STEM_SP -> 64-channel PL view [B,H,W/4]) and level0 as a 64 -> 64
    shift-GEMM over that view with Toeplitz-expanded weights (EPI_SP2SC -> the 16-channel SC map), vs fp64 torch and
    vs the per-pixel SC kernels (dla.py:241-270, :302-312, :325-331)."""
    from sgtapose_b200 import planes as P
    B, (H, W) = 3, hw
    img, hm = C.gen(61, B, 3, H, W), C.gen(62, B, 1, H, W).abs()
    wi, wh = C.gen(63, 16, 3, 7, 7) * 0.1, C.gen(64, 16, 1, 7, 7) * 0.2
    sc, sh = C.gen(65, 32).abs() + 0.5, C.gen(66, 32) * 0.3
    d = lambda t: t.double()
    in4 = P.PlaneBuf(B, 4, H, W, ns, DEV, border=3)
    P.pack_stem(img.to(DEV), hm.to(DEV), in4.full, 0)
    q = torch.cat([img, hm], 1)                                 # the (bf16-)quantised input the kernels see
    if ns == 1:
        q = q.bfloat16().float()
    ref0 = torch.relu(F.conv2d(d(q[:, :3]), d(wi), None, 1, 3) * d(sc)[None, :16, None, None] + d(sh)[None, :16, None, None]) + \
        torch.relu(F.conv2d(d(q[:, 3:]), d(wh), None, 1, 3) * d(sc)[None, 16:, None, None] + d(sh)[None, 16:, None, None])
    spec = P.StemSuperSpec(wi.to(DEV), wh.to(DEV), sc.to(DEV), sh.to(DEV), W, ns)
    f0sp = P.PlaneBuf(B, 64, H, W // 4, ns, DEV)
    P.conv_stem_sp(spec, in4.full, f0sp.full)
    got0 = P.from_superpixels(f0sp.to_nchw(), 4).cpu()
    assert rel_err(got0, ref0) < TOL[ns] * 3
    # the per-pixel stem agrees too (same arithmetic, different tiling)
    stem = P.ScConvSpec([(wi.to(DEV), 0), (wh.to(DEV), 3)], sc.to(DEV), sh.to(DEV), 4, 7, 1, 3, 3, W, ns)
    f0 = P.PlaneBuf(B, 16, H, W, ns, DEV)
    P.conv_sc(stem, in4.full, f0.full, P.EPI_STEM)
    assert rel_err(f0.to_nchw().cpu(), got0) < TOL[ns] * 3
    # level0 over super-pixels -> SC map
    w0 = C.gen(67, 16, 16, 3, 3) * 0.1
    s0, t0 = C.gen(68, 16).abs() + 0.5, C.gen(69, 16) * 0.1
    ref1 = torch.relu(F.conv2d(d(got0), d(w0), None, 1, 1) * d(s0)[None, :, None, None] + d(t0)[None, :, None, None])
    l0spec = P.ConvSpec(P.weight_matrix(P.superpixel_weight(w0.to(DEV), 4, 4, 1)), s0.to(DEV).repeat(4), t0.to(DEV).repeat(4),
                        64, 3, 1, ns, P.ACT_RELU)
    l0 = P.PlaneBuf(B, 16, H, W, ns, DEV)
    P.conv(l0spec, f0sp.full, y=l0.full, epi=P.EPI_SP2SC)
    got1 = l0.to_nchw().cpu()
    assert rel_err(got1, ref1) < TOL[ns] * 3
    assert float(l0.t.view(torch.int16).abs().sum()) > 0


DCN_CFGS = [  # B, Cin, Cout, H, W
    (2, 64, 64, 24, 24),
    (1, 128, 64, 17, 13),
    (1, 64, 128, 12, 20),
    (1, 256, 256, 12, 12),
    (1, 512, 256, 6, 6),
    (3, 128, 128, 16, 16),      # 8 x 16 tiles: dcn_tile_kernel (window staged in shared memory), big offsets -> fallback rows
    (2, 64, 64, 32, 32),        # dcn_tile_kernel, halo 3
    (1, 128, 64, 48, 48),       # BASELINE #6/#8/#10/#12
    (2, 64, 64, 96, 96),        # BASELINE #7/#9/#11/#13/#15
    (1, 64, 128, 8, 48),
]


@pytest.mark.parametrize("cfg", DCN_CFGS)
@pytest.mark.parametrize("ns", [1, 2])
@pytest.mark.parametrize("route", ["gather", "tile"])
def test_dcn_planes_vs_oracle(cfg, ns, route):
    """DeformConv = offset/mask conv (shift-GEMM, fp32 rows out) + fused bilinear-gather DCN GEMM
    with folded bias/BN + ReLU, vs the CPU oracle (torchvision deform_conv2d semantics).  Both producers are
    covered in both modes: the __ldg gather kernel (debug flag 32 = never stage) and the window-staged tile kernel
    (flag 256 = stage whenever the map tiles into 8 x 16 blocks; the default picks per mode)."""
    from sgtapose_b200 import _lib, planes as P
    old = _lib.load().sgta_debug_flags(32 if route == "gather" else 256)
    try:
        _dcn_planes_case(cfg, ns, P)
    finally:
        _lib.load().sgta_debug_flags(old)


def _dcn_planes_case(cfg, ns, P):
    B, Ci, Co, H, W = cfg
    big = Ci == 128
    x = C.gen(21, B, Ci, H, W)
    w = C.gen(22, Co, Ci, 3, 3) * (1.0 / (Ci * 9)) ** 0.5
    omw = C.gen(24, 27, Ci, 3, 3) * (0.3 if big else 0.03)
    omb = C.gen(25, 27)
    scale, shift = C.gen(26, Co).abs() + 0.5, C.gen(27, Co) * 0.2
    xb = _buf(x, ns)
    xq = xb.to_nchw().cpu()                                  # what the kernel actually sees
    om_ref = F.conv2d(xq, omw, omb, padding=1)
    off, mask = odcn.split_offset_mask(om_ref)
    acc = odcn.dcn_v2_conv(xq, off, mask, w, None)
    ref = torch.relu(acc * scale[None, :, None, None] + shift[None, :, None, None])
    omspec = P.ConvSpec(P.weight_matrix(omw.to(DEV)), torch.ones(27, device=DEV), omb.to(DEV), Ci, 3, 1, ns)
    Pn = B * (H + 2) * (W + 2)
    om = torch.zeros(Pn + 256, 32, device=DEV)
    P.conv(omspec, xb.full, y_f32=om, ld_f32=32, epi=P.EPI_F32ROWS)
    wspec = P.ConvSpec(P.weight_matrix(w.to(DEV)), scale.to(DEV), shift.to(DEV), Ci, 3, 1, ns)
    yb = P.PlaneBuf(B, Co, H, W, ns, DEV)
    P.dcn(xb.full, om, wspec, wspec.scale, wspec.shift, yb.full, relu=True)
    err = rel_err(yb.to_nchw().cpu(), ref)
    # bf16 mode: the offsets themselves carry bf16 rounding of the offset conv -> looser bound
    assert err < (1e-4 if ns == 2 else 6e-2), err


@pytest.mark.parametrize("ns", [1, 2])
def test_planes_upsample_and_tokens(ns):
    from sgtapose_b200 import planes as P
    B, Cc, H = 2, 64, 12
    x = C.gen(81, B, Cc, H, H)
    xb = _buf(x, ns)
    xq = xb.to_nchw().cpu()
    for f in (2, 4):
        w = C.gen(82 + f, Cc, 1, 2 * f, 2 * f).abs()
        skip = C.gen(90 + f, B, Cc, H * f, H * f)
        sb = _buf(skip, ns)
        ref = F.conv_transpose2d(xq, w, None, stride=f, padding=f // 2, groups=Cc) + sb.to_nchw().cpu()
        yb = P.PlaneBuf(B, Cc, H * f, H * f, ns, DEV)
        P.upsample_add(xb.full, w.to(DEV), sb.full, yb.full, Cc, f)
        assert rel_err(yb.to_nchw().cpu(), ref) < (4e-3 if ns == 1 else 1e-6)
        if f == 2:          # the four-outputs-per-thread kernel == the general kernel (debug flag 65536), bit for bit
            from sgtapose_b200 import _lib
            y2 = P.PlaneBuf(B, Cc, H * f, H * f, ns, DEV)
            old_flags = _lib.load().sgta_debug_flags(65536)
            try:
                P.upsample_add(xb.full, w.to(DEV), sb.full, y2.full, Cc, f)
            finally:
                _lib.load().sgta_debug_flags(old_flags)
            assert torch.equal(yb.t, y2.t)
    # tokens: gather from the second batch half, deterministic write-back with duplicates
    big = P.PlaneBuf(2 * B, Cc, H, H, ns, DEV)
    big.view(B, B).from_nchw(x.to(DEV))
    ids = torch.tensor([[5, 3, 5, 143, 3, 3, 77], [0, 1, 2, 2, 2, 100, 0]], device=DEV)
    rows = P.gather_tokens(big.full, B, ids, Cc).cpu()
    flat = xq.reshape(B, Cc, H * H).permute(0, 2, 1)
    want = torch.stack([flat[b][ids[b].cpu()] for b in range(B)])
    assert torch.equal(rows, want)
    new = C.gen(99, B, 7, Cc)
    P.scatter_tokens(big.full, B, ids, new.to(DEV), Cc)
    after = big.view(B, B).to_nchw().cpu().reshape(B, Cc, H * H).permute(0, 2, 1)
    exp = flat.clone()
    for b in range(B):
        for t in range(7):
            exp[b, ids[b, t].item()] = new[b, t]
    assert rel_err(after, exp) < (4e-3 if ns == 1 else 3e-7)
    assert float(big.view(0, B).to_nchw().abs().max()) == 0.0


def test_conv_heads_fused_vs_two_stage_and_torch():
    """sgta_planes_conv_heads (the 1x1 output convolutions of the heads in the epilogue of the stacked 3x3 head
    convolution; base_model.py:121-135, :190-199, detector sigmoid sgta_detector.py:854-862) vs the two-stage planes
    path it replaces and vs float64 torch, on a ragged map (the last M tile is partial) with three heads (7 / 2 / 2)."""
    from sgtapose_b200 import planes as P
    B, Cin, H, W, hid = 3, 64, 20, 28, 256
    nouts, sig = [7, 2, 2], [True, False, False]
    x = C.gen(301, B, Cin, H, W)
    xb = _buf(x, 2)
    xq = xb.to_nchw().cpu().double()
    w0 = [C.gen(310 + h, hid, Cin, 3, 3) * 0.05 for h in range(3)]
    b0 = [C.gen(320 + h, hid) * 0.1 for h in range(3)]
    w2 = [C.gen(330 + h, nouts[h], hid, 1, 1) * 0.1 for h in range(3)]
    b2 = [C.gen(340 + h, nouts[h]) * 0.1 for h in range(3)]
    wm = torch.cat([P.weight_matrix(w.to(DEV)) for w in w0], 0)
    bias0 = torch.cat(b0).to(DEV)
    spec0 = P.ConvSpec(wm, torch.ones_like(bias0), bias0, Cin, 3, 1, 2, P.ACT_RELU)
    heads = P.HeadsSpec([w.to(DEV) for w in w2], [b.to(DEV) for b in b2], sig)
    outs = [torch.full((B, n, H, W), float("nan"), device=DEV) for n in nouts]
    P.conv_heads(spec0, heads, xb.full, outs)
    # two-stage planes path
    hb = P.PlaneBuf(B, 3 * hid, H, W, 2, DEV)
    P.conv(spec0, xb.full, hb.full)
    for h in range(3):
        ref = F.conv2d(F.relu(F.conv2d(xq, w0[h].double(), b0[h].double(), padding=1)), w2[h].double(), b2[h].double())
        if sig[h]:
            ref = torch.sigmoid(ref)
        s2 = P.ConvSpec(P.weight_matrix(w2[h].to(DEV)), torch.ones(nouts[h], device=DEV), b2[h].to(DEV), hid, 1, 1, 2,
                        P.ACT_SIGMOID if sig[h] else P.ACT_NONE, n_valid=nouts[h])
        two = torch.zeros(B, nouts[h], H, W, device=DEV)
        P.conv(s2, hb.view(c0=hid * h, C=hid), y_f32=two, epi=P.EPI_NCHW)
        assert torch.isfinite(outs[h]).all()
        assert rel_err(outs[h].cpu().double(), ref) < 2e-6, (h, rel_err(outs[h].cpu().double(), ref))
        assert rel_err(outs[h].cpu(), two.cpu()) < 4e-6, (h, rel_err(outs[h].cpu(), two.cpu()))
