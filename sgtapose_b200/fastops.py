"""NHWC fast-path operators (tcgen05 / TMEM kernels) used by the inference engine.

These are the B200-native forms of the operators behind the reference modules; the
reference-layout (NCHW fp32) module API lives in dcn_v2.py / networks.py.
"""
import torch

from . import _lib

MMA_BF16, MMA_F32X3 = 0, 1
DT_F32, DT_BF16 = 0, 1


def dcn_fast_supported(Cin, Cout):
    return Cin % 64 == 0 and Cout % 32 == 0 and 32 <= Cout <= 256


def pack_dcn_weight(weight, mode):
    """weight [Cout,Cin,3,3] fp32 (CUDA) -> opaque packed buffer for sgta_dcn_forward_nhwc."""
    Cout, Cin, kh, kw = weight.shape
    assert (kh, kw) == (3, 3)
    nbytes = _lib.load().sgta_dcn_wpack_bytes(Cin, Cout, mode)
    if nbytes <= 0:
        raise _lib.SgtaError("DCN fast path does not support Cin=%d Cout=%d" % (Cin, Cout))
    w = weight.detach().contiguous().float()
    buf = torch.empty(nbytes, device=weight.device, dtype=torch.uint8)
    _lib.call("sgta_dcn_pack_weight", _lib.ptr(w), _lib.ptr(buf), Cin, Cout, mode, _lib.stream())
    return buf


def dcn_nhwc(x, om, wpack, scale, shift, Cout, mode, relu=False, out_dtype=None):
    """x [B,H,W,Cin] (bf16 for MMA_BF16, fp32 for MMA_F32X3), om [B,H,W,32] fp32 raw
    conv_offset_mask output (27 used) -> y [B,H,W,Cout] = act(acc*scale + shift)."""
    B, H, W, Cin = x.shape
    want = torch.bfloat16 if mode == MMA_BF16 else torch.float32
    if x.dtype != want:
        raise _lib.SgtaError("dcn_nhwc: mode %d needs %s input" % (mode, want))
    if om.shape != (B, H, W, 32) or om.dtype != torch.float32:
        raise _lib.SgtaError("dcn_nhwc: offset/mask must be [B,H,W,32] fp32")
    out_dtype = out_dtype or x.dtype
    y = torch.empty(B, H, W, Cout, device=x.device, dtype=out_dtype)
    _lib.call("sgta_dcn_forward_nhwc", _lib.ptr(x), _lib.ptr(om), _lib.ptr(wpack), _lib.ptr(scale),
              _lib.ptr(shift), _lib.ptr(y), B, Cin, Cout, H, W, mode, int(relu),
              DT_BF16 if out_dtype == torch.bfloat16 else DT_F32, _lib.stream())
    return y
