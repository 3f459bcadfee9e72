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
