"""Oracle: modulated deformable convolution v2.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The operator is constructed at reference ``sgtapose/lib/model/networks/dla.py:545``
and called at ``dla.py:548``; its arithmetic lives in the third-party, un-vendored
lbin/DCNv2 extension (``README.md:21-28``, branch ``pytorch_<torch version>``, no
pinned commit).  This file restates the published DCNv2 algorithm
(Zhu et al., "Deformable ConvNets v2"): ``modulated_deformable_im2col`` followed by
a GEMM with the flattened weight, as upstream ``dcn_v2_cuda.cu`` /
``dcn_v2_im2col_cuda.cu`` do, and is pinned against
``torchvision.ops.deform_conv2d`` (tests/test_oracle_dcn.py), the CPU stand-in
BASELINE.json names.  Parity with upstream DCNv2 sources themselves is UNPINNED
(they are not in /root/reference).
"""
import torch
import torch.nn.functional as F


def split_offset_mask(om, kh=3, kw=3, deformable_groups=1):
    """``DCN.forward`` upstream: ``o1, o2, mask = chunk(out, 3); offset = cat(o1, o2)``
    which is the identity on the first 2*dg*kh*kw channels; mask = sigmoid(rest).
    Offsets are interleaved (dy, dx) per tap, tap = i*kw + j."""
    n_off = 2 * deformable_groups * kh * kw
    return om[:, :n_off], torch.sigmoid(om[:, n_off:])


def bilinear_zero(img, y, x):
    """DCNv2 ``dmcn_im2col_bilinear`` with its caller's guard: the sample is 0
    unless -1 < y < H and -1 < x < W; any corner outside [0,H)x[0,W) contributes 0.

    img: [B, C, H, W]; y, x: [B, P] (float) -> [B, C, P]
    """
    B, C, H, W = img.shape
    inside = (y > -1) & (y < H) & (x > -1) & (x < W)
    y0 = torch.floor(y)
    x0 = torch.floor(x)
    ly, lx = y - y0, x - x0
    hy, hx = 1 - ly, 1 - lx
    y0 = y0.long()
    x0 = x0.long()
    flat = img.reshape(B, C, H * W)
    out = torch.zeros(B, C, y.shape[1], dtype=img.dtype)
    for dy, dx, w in ((0, 0, hy * hx), (0, 1, hy * lx), (1, 0, ly * hx), (1, 1, ly * lx)):
        yy, xx = y0 + dy, x0 + dx
        ok = inside & (yy >= 0) & (yy < H) & (xx >= 0) & (xx < W)
        idx = (yy.clamp(0, H - 1) * W + xx.clamp(0, W - 1))
        v = torch.gather(flat, 2, idx[:, None, :].expand(B, C, -1))
        out = out + v * (w * ok.to(img.dtype))[:, None, :]
    return out


def deform_im2col(x, offset, mask, kh=3, kw=3, stride=1, pad=1, dil=1, deformable_groups=1):
    """Column matrix [B, Cin*kh*kw (c-major, tap-minor), Ho*Wo] — the layout the
    flattened weight ``weight.view(Cout, Cin*kh*kw)`` contracts against."""
    B, C, H, W = x.shape
    Ho = (H + 2 * pad - (dil * (kh - 1) + 1)) // stride + 1
    Wo = (W + 2 * pad - (dil * (kw - 1) + 1)) // stride + 1
    dg = deformable_groups
    cpg = C // dg
    py = torch.arange(Ho, dtype=x.dtype)[:, None].expand(Ho, Wo).reshape(-1)
    px = torch.arange(Wo, dtype=x.dtype)[None, :].expand(Ho, Wo).reshape(-1)
    cols = torch.zeros(B, C, kh * kw, Ho * Wo, dtype=x.dtype)
    off = offset.reshape(B, dg, kh * kw, 2, Ho * Wo)
    msk = mask.reshape(B, dg, kh * kw, Ho * Wo)
    for g in range(dg):
        xg = x[:, g * cpg:(g + 1) * cpg]
        for i in range(kh):
            for j in range(kw):
                k = i * kw + j
                y = py[None] * stride - pad + i * dil + off[:, g, k, 0]
                xx = px[None] * stride - pad + j * dil + off[:, g, k, 1]
                cols[:, g * cpg:(g + 1) * cpg, k] = bilinear_zero(xg, y, xx) * msk[:, g, k][:, None, :]
    return cols.reshape(B, C * kh * kw, Ho * Wo), Ho, Wo


def dcn_v2_conv(x, offset, mask, weight, bias, stride=1, pad=1, dil=1, deformable_groups=1):
    """out[b,o,p] = bias[o] + sum_{c,k} W[o,c,k] * mask[b,k,p] * bilinear(x[b,c], p + tap_k + d[b,k,p])."""
    Cout, Cin, kh, kw = weight.shape
    cols, Ho, Wo = deform_im2col(x, offset, mask, kh, kw, stride, pad, dil, deformable_groups)
    out = torch.einsum("ok,bkp->bop", weight.reshape(Cout, -1), cols)
    if bias is not None:
        out = out + bias[None, :, None]
    return out.reshape(x.shape[0], Cout, Ho, Wo)


def dcn_forward(x, weight, bias, om_weight, om_bias, stride=1, pad=1, dil=1, deformable_groups=1):
    """Whole ``DCN.forward`` as called at dla.py:548: conv_offset_mask then the
    modulated deformable conv."""
    kh, kw = weight.shape[2:]
    om = F.conv2d(x, om_weight, om_bias, stride=stride, padding=pad)
    offset, mask = split_offset_mask(om, kh, kw, deformable_groups)
    return dcn_v2_conv(x, offset, mask, weight, bias, stride, pad, dil, deformable_groups)


def dcn_forward_torchvision(x, weight, bias, om_weight, om_bias, stride=1, pad=1, dil=1,
                            deformable_groups=1):
    """Same operator through ``torchvision.ops.deform_conv2d`` (multi-threaded C++);
    used for the timed CPU baseline and full-size oracle runs."""
    from torchvision.ops import deform_conv2d
    kh, kw = weight.shape[2:]
    om = F.conv2d(x, om_weight, om_bias, stride=stride, padding=pad)
    n_off = 2 * deformable_groups * kh * kw
    return deform_conv2d(x, om[:, :n_off], weight, bias, stride=stride, padding=pad,
                         dilation=dil, mask=torch.sigmoid(om[:, n_off:]))
