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
