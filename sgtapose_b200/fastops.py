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


# --------------------------------------------------------------------------------------
# plain convolutions / elementwise helpers (conv_umma.cu, elementwise.cu)
# --------------------------------------------------------------------------------------
ACT_NONE, ACT_RELU, ACT_SIGMOID = 0, 1, 2
EPI_NHWC, EPI_STEM, EPI_NCHW = 0, 1, 2


def _esize(dtype):
    return 2 if dtype == torch.bfloat16 else 4


def _dt(dtype):
    return DT_BF16 if dtype == torch.bfloat16 else DT_F32


def pad_to(v, m):
    return (v + m - 1) // m * m


def weight_matrix(weight, cin_pad=None, cin_off=0):
    """conv weight [Cout,Cin,kh,kw] -> Wm [Cout, kh*kw*cin_pad] with K = (tap, channel),
    the weight's channels placed at [cin_off, cin_off+Cin) of each tap."""
    Cout, Cin, kh, kw = weight.shape
    cin_pad = cin_pad or Cin
    wm = weight.new_zeros(Cout, kh * kw, cin_pad)
    wm[:, :, cin_off:cin_off + Cin] = weight.permute(0, 2, 3, 1).reshape(Cout, kh * kw, Cin)
    return wm.reshape(Cout, kh * kw * cin_pad)


class ConvSpec:
    """Packed weights + folded scale/shift of one convolution for sgta_conv_forward_nhwc."""

    def __init__(self, wm, scale, shift, Cin, kh, kw, stride, pad, mode, act=ACT_NONE, n_valid=None):
        dev = wm.device
        Cout, K = wm.shape
        self.n_valid = n_valid or Cout
        Cp, Kp = pad_to(Cout, 16), pad_to(K, 64)
        wmp = torch.zeros(Cp, Kp, device=dev, dtype=torch.float32)
        wmp[:Cout, :K] = wm.float()
        self.scale = torch.ones(Cp, device=dev, dtype=torch.float32)
        self.shift = torch.zeros(Cp, device=dev, dtype=torch.float32)
        self.scale[:Cout] = scale.float()
        self.shift[:Cout] = shift.float()
        nbytes = _lib.load().sgta_conv_wpack_bytes(Cp, Kp, mode)
        if nbytes <= 0:
            raise _lib.SgtaError("conv fast path does not support Cout=%d Kpad=%d" % (Cp, Kp))
        self.wpack = torch.empty(nbytes, device=dev, dtype=torch.uint8)
        _lib.call("sgta_conv_pack_weight", _lib.ptr(wmp), _lib.ptr(self.wpack), Cp, Kp, mode, _lib.stream())
        self.Cin, self.Cout, self.kh, self.kw = Cin, Cp, kh, kw
        self.stride, self.pad, self.mode, self.act = stride, pad, mode, act

    def out_hw(self, H, W):
        return ((H + 2 * self.pad - self.kh) // self.stride + 1, (W + 2 * self.pad - self.kw) // self.stride + 1)


def _ptr_off(t, coff=0):
    return t.data_ptr() + coff * t.element_size()


def conv_nhwc(spec, x, B, H, W, ldx, y, ldy, x_coff=0, y_coff=0, res=None, ldres=0, res_coff=0,
              epi=EPI_NHWC, act=None):
    """x: tensor holding [B,H,W,ldx] (channels [x_coff, x_coff+Cin) are read); y likewise."""
    want = torch.bfloat16 if spec.mode == MMA_BF16 else torch.float32
    if x.dtype != want:
        raise _lib.SgtaError("conv_nhwc: mode %d needs %s input, got %s" % (spec.mode, want, x.dtype))
    _lib.call("sgta_conv_forward_nhwc", _ptr_off(x, x_coff), ldx, _lib.ptr(spec.wpack), _lib.ptr(spec.scale),
              _lib.ptr(spec.shift), _ptr_off(res, res_coff) if res is not None else None, ldres,
              _ptr_off(y, y_coff), ldy, B, H, W, spec.Cin, spec.Cout, spec.kh, spec.kw, spec.stride,
              spec.pad, spec.mode, spec.act if act is None else act, _dt(y.dtype),
              _dt(res.dtype) if res is not None else 0, epi, spec.n_valid, _lib.stream())
    return y


def nchw_to_nhwc(src, dst, ld, coff=0):
    B, C, H, W = src.shape
    _lib.call("sgta_nchw_to_nhwc", _lib.ptr(src), dst.data_ptr(), B, C, H * W, ld, coff, _dt(dst.dtype),
              _lib.stream())


def nhwc_to_nchw(src, dst, C, ld, coff=0):
    B, _, H, W = dst.shape
    _lib.call("sgta_nhwc_to_nchw", src.data_ptr(), _lib.ptr(dst), B, C, H * W, ld, coff, _dt(src.dtype),
              _lib.stream())


def maxpool2(x, B, H, W, C, ldx, y, ldy, x_coff=0, y_coff=0):
    _lib.call("sgta_maxpool2x2_nhwc", _ptr_off(x, x_coff), ldx, _ptr_off(y, y_coff), ldy, B, H, W, C,
              _dt(x.dtype), _lib.stream())


def upsample_add(x, w_up, skip, ldskip, y, ldy, B, h, w, C, f, skip_coff=0):
    _lib.call("sgta_upsample_add_nhwc", _lib.ptr(x), _lib.ptr(w_up),
              _ptr_off(skip, skip_coff) if skip is not None else None, ldskip, _lib.ptr(y), ldy, B, h, w, C, f,
              _dt(x.dtype), _lib.stream())


def gather_tokens_nhwc(feats, ld, ids, C, HW):
    B, n = ids.shape
    rows = torch.empty(B, n, C, device=ids.device, dtype=torch.float32)
    _lib.call("sgta_gather_tokens_nhwc", feats.data_ptr(), ld, _lib.ptr(ids), _lib.ptr(rows), B, C, HW, n,
              _dt(feats.dtype), _lib.stream())
    return rows


def scatter_tokens_nhwc(feats, ld, ids, rows, C, HW):
    B, n = ids.shape
    rows = rows.contiguous().float()
    _lib.call("sgta_scatter_tokens_nhwc", feats.data_ptr(), ld, _lib.ptr(ids), _lib.ptr(rows), B, C, HW, n,
              _dt(feats.dtype), _lib.stream())
