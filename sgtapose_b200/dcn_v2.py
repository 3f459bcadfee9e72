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

Inference route (autograd off, 3x3 / stride 1 / pad 1 / dilation 1 / one deformable group, Cin and Cout multiples
of 64 -- every DCN of the hot path, dla.py:545): the same tcgen05 kernels as the compiled engine.  The NCHW input
is packed once into the fp16 hi/lo planes layout (csrc/planes.cuh), conv_offset_mask runs as a shift-GEMM
(`sgta_planes_conv`), the bilinear gather + implicit GEMM as `sgta_planes_dcn`, and the result is unpacked to NCHW
fp32.  Everything else (training, other geometries) takes the exact-fp32 SIMT kernels.  `DCN.tensor_core = False`
turns the route off.
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

    tensor_core = True        # class-wide switch for the inference route

    def _tc_eligible(self, x):
        return (self.tensor_core and x.is_cuda and not torch.is_grad_enabled() and x.dim() == 4
                and self.kernel_size == (3, 3) and self.stride == 1 and self.padding == 1 and self.dilation == 1
                and self.deformable_groups == 1 and self.in_channels % 64 == 0 and self.out_channels % 64 == 0)

    def _tc_forward(self, x):
        """tcgen05 route: planes in, planes out, same kernels as engine.InferenceEngine._deform_conv."""
        from . import planes as P
        dev = x.device
        B, Cin, H, W = x.shape
        Cout = self.out_channels
        params = (self.weight, self.bias, self.conv_offset_mask.weight, self.conv_offset_mask.bias)
        wkey = tuple((t.data_ptr(), t._version) for t in params) + (str(dev),)
        cache = self.__dict__.setdefault("_tc_cache", {})
        if cache.get("wkey") != wkey:
            ones = lambda n: torch.ones(n, device=dev)
            cache["om"] = P.ConvSpec(P.weight_matrix(self.conv_offset_mask.weight.detach().float()), ones(27),
                                     self.conv_offset_mask.bias.detach().float(), Cin, 3, 1, 2)       # N 27 -> 32
            cache["w"] = P.ConvSpec(P.weight_matrix(self.weight.detach().float()), ones(Cout),
                                    self.bias.detach().float(), Cin, 3, 1, 2)
            cache["wkey"] = wkey
        skey = (B, H, W)
        if cache.get("skey") != skey:
            cache["x"] = P.PlaneBuf(B, Cin, H, W, 2, dev)
            cache["y"] = P.PlaneBuf(B, Cout, H, W, 2, dev)
            cache["omf"] = torch.zeros(B * (H + 2) * (W + 2) + 256, 32, device=dev, dtype=torch.float32)
            cache["skey"] = skey
        xb, yb, w = cache["x"], cache["y"], cache["w"]
        xb.from_nchw(x)
        P.conv(cache["om"], xb.full, y_f32=cache["omf"], ld_f32=32, epi=P.EPI_F32ROWS)
        P.dcn(xb.full, cache["omf"], w, w.scale, w.shift, yb.full, relu=False)
        return yb.to_nchw()

    def forward(self, x):
        if self._tc_eligible(x):
            return self._tc_forward(x)
        om = self.conv_offset_mask(x)
        return dcn_v2_conv(x, om, self.weight, self.bias, self.stride, self.padding, self.dilation,
                           self.deformable_groups)

    def extra_repr(self):
        return "%d, %d, kernel_size=%s, stride=%d, padding=%d, dilation=%d, deformable_groups=%d" % (
            self.in_channels, self.out_channels, self.kernel_size, self.stride, self.padding,
            self.dilation, self.deformable_groups)
