"""Host-side helpers of the planes layout that need no GPU."""
import pytest
import torch
import torch.nn.functional as F


@pytest.mark.parametrize("Ci,Co,stride,gin,gout,W", [(16, 16, 1, 4, 4, 16), (16, 32, 2, 4, 4, 32), (16, 32, 2, 4, 2, 16),
                                                       (32, 64, 2, 2, 1, 16), (4, 8, 1, 2, 2, 10)])
def test_superpixel_weight_is_the_same_convolution(Ci, Co, stride, gin, gout, W):
    """A 3x3 convolution over super-pixels with the expanded weights == the original convolution."""
    from sgtapose_b200 import planes
    torch.manual_seed(1)
    x = torch.randn(2, Ci, 12, W, dtype=torch.float64)
    w = torch.randn(Co, Ci, 3, 3, dtype=torch.float64)
    ref = F.conv2d(x, w, stride=stride, padding=1)
    wp = planes.superpixel_weight(w, gin, gout, stride)
    assert wp.shape == (gout * Co, gin * Ci, 3, 3)
    y = F.conv2d(planes.to_superpixels(x, gin), wp, stride=(stride, stride * gout // gin), padding=1)
    assert torch.allclose(planes.from_superpixels(y, gout), ref, atol=1e-12, rtol=0)
    assert torch.equal(planes.from_superpixels(planes.to_superpixels(x, gin), gin), x)
    with pytest.raises(ValueError):
        planes.superpixel_weight(w, 3, 2, 1)


def test_load_model_save_model_reference_call_forms(tmp_path):
    """model.py:43-114 call forms: `(model, optimizer, start_epoch)` with an optimizer (lr stepped under opt.resume),
    `module.` prefixes stripped, DataParallel unwrapped on save, mismatching `hm` tensors skipped or re-used."""
    import types
    import torch
    from torch import nn
    from sgtapose_b200 import networks

    class Tiny(nn.Module):
        def __init__(self, n_hm):
            super().__init__()
            self.body = nn.Linear(4, 4)
            self.hm = nn.Linear(4, n_hm)

    src = Tiny(7)
    path = str(tmp_path / "m.pth")
    opt_src = torch.optim.SGD(src.parameters(), lr=1.0)
    networks.save_model(path, 25, nn.DataParallel(src), opt_src)
    saved = torch.load(path)
    assert saved["epoch"] == 25 and "optimizer" in saved and all(not k.startswith("module") for k in saved["state_dict"])
    dst = Tiny(7)
    assert networks.load_model(dst, path, types.SimpleNamespace()) is dst
    assert all(torch.equal(a, b) for a, b in zip(src.state_dict().values(), dst.state_dict().values()))
    opt = types.SimpleNamespace(resume=True, lr=1e-3, lr_step=[10, 20, 30], reset_hm=False, reuse_hm=False)
    o = torch.optim.SGD(dst.parameters(), lr=1e-3)
    m, o2, ep = networks.load_model(dst, path, opt, o)
    assert m is dst and o2 is o and ep == 25 and abs(o.param_groups[0]["lr"] - 1e-5) < 1e-12
    small = Tiny(3)
    keep = small.hm.weight.clone()
    networks.load_model(small, path, types.SimpleNamespace(reset_hm=False, reuse_hm=False))
    assert torch.equal(small.hm.weight, keep) and torch.equal(small.body.weight, src.body.weight)
    networks.load_model(small, path, types.SimpleNamespace(reset_hm=False, reuse_hm=True))
    assert torch.equal(small.hm.weight, src.hm.weight[:3])


def test_stem_superpixel_matrix_vs_conv2d():
    """Host half of the super-pixel stem (planes.stem_superpixel_matrix): the Toeplitz-expanded [128, 7*64] matrix,
    applied to K blocks gathered exactly like the SC gather kernel does (kernel row ky = 16 padded input pixels x 4
    channels from the group's first column, 8-byte-pixel order), equals the two 7x7 convolutions of dla.py:241-270."""
    import torch
    import torch.nn.functional as F
    from sgtapose_b200 import planes as P
    g = torch.Generator().manual_seed(5)
    B, H, W = 2, 6, 12
    img, hm = torch.randn(B, 3, H, W, generator=g, dtype=torch.float64), torch.randn(B, 1, H, W, generator=g, dtype=torch.float64)
    wi, wh = torch.randn(16, 3, 7, 7, generator=g, dtype=torch.float64), torch.randn(16, 1, 7, 7, generator=g, dtype=torch.float64)
    wm = P.stem_superpixel_matrix(wi, wh)
    assert wm.shape == (128, 448)
    x = F.pad(torch.cat([img, hm], 1), (3, 3 + 16, 3, 3))            # border 3 (+ slack on the right for the 16-pixel runs)
    ref = torch.cat([F.conv2d(img, wi, None, 1, 3), F.conv2d(hm, wh, None, 1, 3)], 1)      # [B,32,H,W]
    for b in range(B):
        for y in range(H):
            for X in range(W // 4):
                a = torch.zeros(448, dtype=torch.float64)
                for ky in range(7):
                    for i in range(16):
                        a[ky * 64 + (i % 8) * 8 + (i // 8) * 4:ky * 64 + (i % 8) * 8 + (i // 8) * 4 + 4] = x[b, :, y + ky, 4 * X + i]
                out = wm @ a
                for j in range(4):
                    assert torch.allclose(out[32 * j:32 * j + 32], ref[b, :, y, 4 * X + j], atol=1e-5)     # the matrix is built in fp32


def test_heads_spec_packing_is_the_1x1_convolutions():
    """HeadsSpec (host half of sgta_planes_conv_heads; base_model.py:121-135): the packed rows [hid][stride] + biases
    [n_heads][8] reproduce the three 1x1 output convolutions when applied the way the kernel's epilogue applies them
    (out[o] = sum_k hidden[k] * w2[k][o] + b2[o]), strides 8 / 2 / 2 for 7 / 2 / 2 outputs, zero padding elsewhere."""
    import torch.nn.functional as F
    from sgtapose_b200 import planes as P
    torch.manual_seed(3)
    hid, nouts = 256, [7, 2, 2]
    ws = [torch.randn(n, hid, 1, 1) for n in nouts]
    bs = [torch.randn(n) for n in nouts]
    hs = P.HeadsSpec(ws, bs, [True, False, False])
    assert hs.nout == nouts and hs.hid == hid and hs.sig_mask == 1
    strides = [8, 2, 2]
    assert hs.w2.numel() == hid * sum(strides) and tuple(hs.b2.shape) == (3, 8)
    hidden = torch.relu(torch.randn(5, 3 * hid, 4, 6))
    off = 0
    for h, (n, st) in enumerate(zip(nouts, strides)):
        m = hs.w2[off:off + hid * st].view(hid, st)
        off += hid * st
        assert torch.count_nonzero(m[:, n:]) == 0 and torch.count_nonzero(hs.b2[h, n:]) == 0
        x = hidden[:, h * hid:(h + 1) * hid]
        got = torch.einsum("bkyx,ko->boyx", x.double(), m[:, :n].double()) + hs.b2[h, :n].double().view(1, n, 1, 1)
        ref = F.conv2d(x.double(), ws[h].double(), bs[h].double())
        assert torch.allclose(got, ref, rtol=1e-12, atol=1e-12)
