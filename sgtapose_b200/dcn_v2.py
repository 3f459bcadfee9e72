"""Drop-in `DCN` module: modulated deformable convolution v2 on B200.

Replaces the un-vendored third-party `DCNv2.dcn_v2.DCN` the reference imports at
sgtapose/lib/model/networks/dla.py:21-25 and constructs at dla.py:545 as
`DCN(chi, cho, kernel_size=(3,3), stride=1, padding=1, dilation=1, deformable_groups=1)`.
Same constructor keywords, same parameter names (`weight`, `bias`,
`conv_offset_mask.weight`, `conv_offset_mask.bias`) and the same init as upstream
(weight ~ U(+-1/sqrt(Cin*kh*kw)), bias = 0, conv_offset_mask = 0), so reference
checkpoints load unchanged (model.py:43-84).

forward(x: [B,Cin,H,W] fp32 CUDA) -> [B,Cout,Ho,Wo]; differentiable w.r.t. x, weight, bias
and conv_offset_mask.* through hand-written CUDA kernels (sgta_dcn_forward/backward).
There is no CPU fallback: a non-CUDA input raises.
"""
import math

import torch
from torch import nn

from . import _lib


def _pair(v):
    return (v, v) if isinstance(v, int) else tuple(v)


class _DCNv2Function(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, offset_mask, weight, bias, stride, pad, dil, dg):
        if not x.is_cuda:
            raise _lib.SgtaError("DCN.forward: input must be a CUDA tensor (no CPU fallback)")
        x = x.contiguous().float()
        offset_mask = offset_mask.contiguous().float()
        weight_c = weight.contiguous().float()
        bias_c = bias.contiguous().float() if bias is not None else None
        B, Cin, H, W = x.shape
        Cout, _, kh, kw = weight_c.shape
        Ho = (H + 2 * pad - (dil * (kh - 1) + 1)) // stride + 1
        Wo = (W + 2 * pad - (dil * (kw - 1) + 1)) // stride + 1
        y = torch.empty(B, Cout, Ho, Wo, device=x.device, dtype=torch.float32)
        _lib.call("sgta_dcn_forward", _lib.ptr(x), _lib.ptr(offset_mask), _lib.ptr(weight_c),
                  _lib.ptr(bias_c), _lib.ptr(y), B, Cin, Cout, H, W, kh, kw, stride, pad, dil, dg,
                  0, _lib.stream())
        ctx.save_for_backward(x, offset_mask, weight_c)
        ctx.cfg = (stride, pad, dil, dg, bias is not None)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, om, w = ctx.saved_tensors
        stride, pad, dil, dg, has_bias = ctx.cfg
        gy = gy.contiguous().float()
        B, Cin, H, W = x.shape
        Cout, _, kh, kw = w.shape
        need = ctx.needs_input_grad
        gx = torch.empty_like(x) if need[0] else None
        gom = torch.empty_like(om) if need[1] else None
        gw = torch.zeros_like(w) if need[2] else None
        gb = torch.zeros(Cout, device=x.device, dtype=torch.float32) if (need[3] and has_bias) else None
        if gb is not None and gw is None:
            gw = torch.zeros_like(w)
        _lib.call("sgta_dcn_backward", _lib.ptr(x), _lib.ptr(om), _lib.ptr(w), _lib.ptr(gy),
                  _lib.ptr(gx), _lib.ptr(gom), _lib.ptr(gw), _lib.ptr(gb), B, Cin, Cout, H, W, kh, kw,
                  stride, pad, dil, dg, 0, _lib.stream())
        return gx, gom, (gw if need[2] else None), gb, None, None, None, None


def dcn_v2_conv(x, offset_mask, weight, bias, stride=1, padding=1, dilation=1, deformable_groups=1):
    """Functional form: `offset_mask` is the RAW conv_offset_mask output
    ([B, 3*dg*kh*kw, Ho, Wo]); the mask sigmoid is fused into the kernel."""
    return _DCNv2Function.apply(x, offset_mask, weight, bias, stride, padding, dilation,
                                deformable_groups)


class DCN(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size=(3, 3), stride=1, padding=1,
                 dilation=1, deformable_groups=1):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size = _pair(kernel_size)
        self.stride, self.padding, self.dilation = int(stride), int(padding), int(dilation)
        self.deformable_groups = int(deformable_groups)
        if in_channels % self.deformable_groups:
            raise ValueError("in_channels must be divisible by deformable_groups")
        kh, kw = self.kernel_size
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels, kh, kw))
        self.bias = nn.Parameter(torch.zeros(out_channels))
        self.conv_offset_mask = nn.Conv2d(in_channels, self.deformable_groups * 3 * kh * kw,
                                          kernel_size=self.kernel_size, stride=self.stride,
                                          padding=self.padding, bias=True)
        self.reset_parameters()

    def reset_parameters(self):
        kh, kw = self.kernel_size
        stdv = 1.0 / math.sqrt(self.in_channels * kh * kw)
        with torch.no_grad():
            self.weight.uniform_(-stdv, stdv)
            self.bias.zero_()
            self.conv_offset_mask.weight.zero_()
            self.conv_offset_mask.bias.zero_()

    def forward(self, x):
        om = self.conv_offset_mask(x)
        return dcn_v2_conv(x, om, self.weight, self.bias, self.stride, self.padding, self.dilation,
                           self.deformable_groups)

    def extra_repr(self):
        return "%d, %d, kernel_size=%s, stride=%d, padding=%d, dilation=%d, deformable_groups=%d" % (
            self.in_channels, self.out_channels, self.kernel_size, self.stride, self.padding,
            self.dilation, self.deformable_groups)
