"""Host side of the engine's native activation layouts and the tcgen05 convolutions that
run on them (csrc/planes.cuh, conv_planes.cu, planes_ops.cu; DESIGN.md "Data layout").

A `PlaneBuf` owns one zero-bordered feature map [B, C, H, W] stored as `nplanes` 16-bit
planes (1 = bf16; 2 = fp16 hi + fp16 2^11-scaled residual, i.e. fp32 to ~2^-22):
  layout PL (C % 64 == 0): [plane][C/64][rows][64], SWIZZLE_128B pre-applied per row;
  layout SC (C in {4,16,32}): [plane][rows][C].
`rows` = guard + B*(H+2b)*(W+2b) + guard.  Views (`.view`) select an image range and, for
PL, a 64-channel-aligned channel range without copying (Root concats, batch halves).
"""
import ctypes

import torch

from . import _lib

PL, SC = 0, 1
EPI_PL, EPI_SC, EPI_F32ROWS, EPI_NCHW, EPI_STEM, EPI_STEM_SP, EPI_SP2SC = 0, 1, 2, 3, 4, 5, 6
ACT_NONE, ACT_RELU, ACT_SIGMOID = 0, 1, 2


SgtaPlanes = _lib.SgtaPlanes


def guard_rows(W):
    return int(_lib.load().sgta_planes_guard(W))


class PlaneView:
    """A (batch range, channel range) window of a PlaneBuf; what the C ABI consumes."""

    def __init__(self, buf, b0, B, c0, C):
        self.buf, self.b0, self.B, self.c0, self.C = buf, b0, B, c0, C
        frame = (buf.H + 2 * buf.border) * (buf.W + 2 * buf.border)
        if buf.layout == PL:
            if c0 % 64 or C % 64:
                raise _lib.SgtaError("PL views are 64-channel aligned")
            nch, ch0 = buf.C // 64, c0 // 64
        else:
            if c0 != 0 or C != buf.C:
                raise _lib.SgtaError("SC views cover all channels")
            nch, ch0 = buf.C, 0
        self.c = SgtaPlanes(buf.t.data_ptr(), buf.rows, buf.guard + b0 * frame, nch, ch0, buf.nplanes,
                            buf.layout, buf.border, B, buf.H, buf.W)
        self.ref = ctypes.byref(self.c)

    @property
    def H(self):
        return self.buf.H

    @property
    def W(self):
        return self.buf.W

    def to_nchw(self):
        out = torch.empty(self.B, self.C, self.buf.H, self.buf.W, device=self.buf.t.device, dtype=torch.float32)
        _lib.call("sgta_planes_to_nchw", self.ref, _lib.ptr(out), self.C, 0, _lib.stream())
        return out

    def from_nchw(self, src):
        src = src.contiguous().float()
        assert tuple(src.shape) == (self.B, self.C, self.buf.H, self.buf.W), (src.shape, self.B, self.C)
        _lib.call("sgta_planes_from_nchw", _lib.ptr(src), self.ref, self.C, 0, _lib.stream())
        return self


class PlaneBuf:
    def __init__(self, B, C, H, W, nplanes, device, layout=None, border=1):
        self.B, self.C, self.H, self.W, self.nplanes, self.border = B, C, H, W, nplanes, border
        self.layout = layout if layout is not None else (PL if C % 64 == 0 else SC)
        if self.layout == PL and (C % 64 or border != 1):
            raise _lib.SgtaError("PL buffers need C % 64 == 0 and border 1")
        if self.layout == SC and C not in (4, 16, 32):
            raise _lib.SgtaError("SC buffers hold 4, 16 or 32 channels")
        g = guard_rows(W + 2 * border)
        self.guard = g
        self.rows = g + B * (H + 2 * border) * (W + 2 * border) + g + 128
        self.rows = (self.rows + 7) // 8 * 8
        n = nplanes * self.rows * (C if self.layout == SC else 64 * (C // 64))
        self.t = torch.zeros(n, device=device, dtype=torch.int16)
        self.full = PlaneView(self, 0, B, 0, C)

    def view(self, b0=0, B=None, c0=0, C=None):
        return PlaneView(self, b0, self.B - b0 if B is None else B, c0, self.C - c0 if C is None else C)

    @property
    def ref(self):
        return self.full.ref

    def to_nchw(self):
        return self.full.to_nchw()

    def from_nchw(self, src):
        self.full.from_nchw(src)
        return self


def pad_to(v, m):
    return (v + m - 1) // m * m


def weight_matrix(weight, cin_pad=None, cin_off=0):
    """conv weight [Cout,Cin,kh,kw] -> Wm [Cout, kh*kw*cin_pad], K = (tap, channel)."""
    Cout, Cin, kh, kw = weight.shape
    cin_pad = cin_pad or Cin
    wm = weight.new_zeros(Cout, kh * kw, cin_pad)
    wm[:, :, cin_off:cin_off + Cin] = weight.permute(0, 2, 3, 1).reshape(Cout, kh * kw, Cin)
    return wm.reshape(Cout, kh * kw * cin_pad)


def superpixel_weight(weight, gin, gout, stride=1):
    """Toeplitz expansion of a 3x3 convolution on `[pixels][C]` rows into a 3x3 convolution over SUPER-PIXELS
    (g consecutive pixels x C channels, channel index j*C + c): weight [Co,Ci,3,3] -> [gout*Co, gin*Ci, 3, 3],
    row stride `stride`, super-column stride `stride*gout/gin`, zero border of one super-pixel.
    With gin*Ci = 64 a small-channel layer (level0: 16 -> 16, level1: 16 -> 32) becomes a plain PL convolution that
    conv_shift_kernel / conv_gather_kernel<STRIDE> run as they are (host half of that plan, checked
    against F.conv2d in tests/test_planes_host.py -- the engine does not use it yet)."""
    Co, Ci, kh, kw = weight.shape
    if (kh, kw) != (3, 3) or (stride * gout) % gin:
        raise ValueError("superpixel_weight: 3x3 kernels, stride*gout divisible by gin")
    out = weight.new_zeros(gout * Co, gin * Ci, 3, 3)
    for j_out in range(gout):
        for n in (-1, 0, 1):
            for j_in in range(gin):
                dx = gin * n + j_in - stride * j_out          # input pixel offset this (tap, pixel) pair stands for
                if -1 <= dx <= 1:
                    out[j_out * Co:(j_out + 1) * Co, j_in * Ci:(j_in + 1) * Ci, :, n + 1] = weight[:, :, :, dx + 1]
    return out


def to_superpixels(x, g):
    """[B,C,H,W] -> [B,g*C,H,W/g] (channel j*C + c = pixel j of the group): the NCHW view of the super-pixel layout."""
    B, C, H, W = x.shape
    return x.reshape(B, C, H, W // g, g).permute(0, 4, 1, 2, 3).reshape(B, g * C, H, W // g)


def from_superpixels(x, g):
    B, GC, H, Ws = x.shape
    return x.reshape(B, g, GC // g, H, Ws).permute(0, 2, 3, 4, 1).reshape(B, GC // g, H, Ws * g)


class ConvSpec:
    """Packed weights + folded scale/shift of one convolution on a PL input."""

    def __init__(self, wm, scale, shift, Cin, ksize, stride, nplanes, act=ACT_NONE, n_valid=None):
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
        nbytes = _lib.load().sgta_planes_wpack_bytes(Cp, Kp, nplanes)
        if nbytes <= 0:
            raise _lib.SgtaError("planes conv does not support Cout=%d Kpad=%d" % (Cp, Kp))
        self.wpack = torch.empty(nbytes, device=dev, dtype=torch.uint8)
        _lib.call("sgta_planes_pack_weight", _lib.ptr(wmp), _lib.ptr(self.wpack), Cp, Kp, nplanes, _lib.stream())
        self.Cin, self.Cout, self.ksize, self.stride, self.nplanes, self.act = Cin, Cp, ksize, stride, nplanes, act


def conv(spec, x, y=None, res=None, y_f32=None, ld_f32=0, epi=EPI_PL, act=None):
    """x: PlaneView (PL).  y: PlaneView for EPI_PL; y_f32: fp32 tensor for EPI_F32ROWS / EPI_NCHW."""
    _lib.call("sgta_planes_conv", x.ref, _lib.ptr(spec.wpack), _lib.ptr(spec.scale), _lib.ptr(spec.shift),
              res.ref if res is not None else None, y.ref if y is not None else None,
              _lib.ptr(y_f32) if y_f32 is not None else None, ld_f32, spec.Cin, spec.Cout, spec.ksize, spec.stride,
              spec.act if act is None else act, epi, spec.n_valid, _lib.stream())


class ScConvSpec:
    """A convolution on an SC input (C in {4,16,32}).  One K block = 128 bytes per output
    pixel = `pix_per_seg` consecutive input pixels x C channels per segment; kernel-row taps
    beyond kw get zero weights."""

    def __init__(self, weight_groups, scale, shift, C, ksize, stride, pad, in_border, in_W, nplanes,
                 act=ACT_NONE):
        """weight_groups: list of (weight [Cout_g, Cin_g, k, k], cin_off) stacked along Cout
        (the dual stem stacks the image and heat-map convolutions)."""
        dev = weight_groups[0][0].device
        k = ksize
        Wp = in_W + 2 * in_border
        row_bytes = C * 2
        # a kernel row needs k pixels; it is cut into segments of 64 or 128 bytes
        need = k * row_bytes
        if need <= 64:
            seg_bytes, segs_per_row = 64, 1
        else:
            seg_bytes, segs_per_row = 128, (need + 127) // 128
        pix_per_seg = seg_bytes // row_bytes
        segs = []                       # (ky, first kx) per segment, kernel-row major
        for ky in range(k):
            for s in range(segs_per_row):
                segs.append((ky, s * pix_per_seg))
        per_block = 128 // seg_bytes
        while len(segs) % per_block:
            segs.append((None, 0))       # dummy segment: zero weights, reads a valid row
        nkb = len(segs) // per_block
        if nkb > 8:
            raise _lib.SgtaError("SC conv: too many K blocks")
        off = in_border - pad
        seg_off = []
        for i in range(nkb * 2):
            j = i // 2 * per_block + (i % 2 if per_block == 2 else 0)
            ky, kx0 = segs[j]
            seg_off.append(((ky if ky is not None else 0) + off) * Wp + kx0 + off)
        self.seg_off = (ctypes.c_int * (nkb * 2))(*seg_off)
        self.nkb, self.seg_groups = nkb, seg_bytes // 16
        Cout = sum(w.shape[0] for w, _ in weight_groups)
        wm = torch.zeros(Cout, nkb * 64, device=dev, dtype=torch.float32)
        o0 = 0
        for w, cin_off in weight_groups:
            co, ci = w.shape[:2]
            for j, (ky, kx0) in enumerate(segs):
                if ky is None:
                    continue
                for pp in range(pix_per_seg):
                    kx = kx0 + pp
                    if kx >= k:
                        continue
                    if C == 4:
                        # 8-byte pixels: the kernel interleaves the two segments of a K block per pixel
                        kbase = (j // 2) * 64 + pp * 8 + (j % 2) * 4 + cin_off
                    else:
                        kbase = j * (seg_bytes // 2) + pp * C + cin_off
                    wm[o0:o0 + co, kbase:kbase + ci] = w[:, :, ky, kx].float()
            o0 += co
        Cp = pad_to(Cout, 16)
        wmp = torch.zeros(Cp, nkb * 64, device=dev, dtype=torch.float32)
        wmp[:Cout] = wm
        self.scale = torch.ones(Cp, device=dev, dtype=torch.float32)
        self.shift = torch.zeros(Cp, device=dev, dtype=torch.float32)
        self.scale[:Cout] = scale.float()
        self.shift[:Cout] = shift.float()
        nbytes = _lib.load().sgta_planes_wpack_bytes(Cp, nkb * 64, nplanes)
        self.wpack = torch.empty(nbytes, device=dev, dtype=torch.uint8)
        _lib.call("sgta_planes_pack_weight", _lib.ptr(wmp), _lib.ptr(self.wpack), Cp, nkb * 64, nplanes, _lib.stream())
        self.Cout, self.stride, self.ksize, self.pad, self.act = Cp, stride, ksize, pad, act


class HeadsSpec:
    """The 1x1 output convolutions of the heads, packed for sgta_planes_conv_heads (base_model.py:121-135): per head the
    [nout, hid] weight as [hid][stride] fp32 rows (stride 2 / 4 / 8, zero padded), heads back to back; biases [n_heads][8]."""

    def __init__(self, weights, biases, sigmoid):
        """weights: list of [nout, hid(,1,1)] tensors; biases: list of [nout]; sigmoid: list of bool."""
        dev = weights[0].device
        self.hid = weights[0].reshape(weights[0].shape[0], -1).shape[1]
        self.nout = [int(w.shape[0]) for w in weights]
        parts = []
        self.b2 = torch.zeros(len(weights), 8, device=dev, dtype=torch.float32)
        for h, (w, b) in enumerate(zip(weights, biases)):
            no = self.nout[h]
            if no > 8 or w.reshape(no, -1).shape[1] != self.hid:
                raise ValueError("HeadsSpec: heads of up to 8 outputs over the same hidden width")
            stride = 2 if no <= 2 else 4 if no <= 4 else 8
            m = torch.zeros(self.hid, stride, device=dev, dtype=torch.float32)
            m[:, :no] = w.reshape(no, self.hid).t().float()
            parts.append(m.reshape(-1))
            self.b2[h, :no] = b.float()
        self.w2 = torch.cat(parts).contiguous()
        self.sig_mask = sum(1 << h for h, s in enumerate(sigmoid) if s)
        self._nout_c = (ctypes.c_int * len(weights))(*self.nout)


def conv_heads(spec, heads, x, outs):
    """spec: ConvSpec of the stacked 3x3 head convolutions (Cin -> n_heads * hid, ReLU); heads: HeadsSpec;
    outs: list of contiguous fp32 [B, nout_h, H, W] tensors.  One launch; fp32 mode only."""
    arr = (ctypes.c_void_p * len(outs))(*[_lib.ptr(o) for o in outs])
    _lib.call("sgta_planes_conv_heads", x.ref, _lib.ptr(spec.wpack), _lib.ptr(spec.scale), _lib.ptr(spec.shift),
              _lib.ptr(heads.w2), _lib.ptr(heads.b2), arr, heads._nout_c, len(outs), heads.sig_mask, spec.Cin, heads.hid,
              _lib.stream())


def conv_sc(spec, x, y, epi, act=None):
    Ho = (x.H + 2 * spec.pad - spec.ksize) // spec.stride + 1
    Wo = (x.W + 2 * spec.pad - spec.ksize) // spec.stride + 1
    _lib.call("sgta_planes_conv_sc", x.ref, _lib.ptr(spec.wpack), _lib.ptr(spec.scale), _lib.ptr(spec.shift), y.ref,
              spec.Cout, spec.stride, spec.stride, Ho, Wo, spec.nkb, spec.seg_groups, spec.seg_off,
              spec.act if act is None else act, epi, _lib.stream())


SP_G = 4        # pixels per super-pixel


def stem_superpixel_matrix(wi, wh):
    """Weight matrix [128, 7*64] of the dual 7x7 stem (dla.py:241-270, :325-331) over SUPER-PIXELS of 4 output pixels:
    column block j*32 = output pixel j of the group ([16 image-conv | 16 heat-map-conv] channels); K block ky = kernel
    row ky = the 16 input pixels x 4 channels [img(3) | hm(1)] starting at the group's first padded column, in the
    8-byte-pixel order of the SC gather (K index = (i % 8)*8 + (i // 8)*4 + ch for input pixel i).  Output pixel j
    reads input pixel i with kernel column kx = i - j (Toeplitz expansion; 10 of the 16 pixels carry weights)."""
    wm = wi.new_zeros(SP_G * 32, 7 * 64)
    for j in range(SP_G):
        for w, cin_off, n_off in ((wi, 0, 0), (wh, 3, 16)):
            co, ci = w.shape[:2]
            for ky in range(7):
                for i in range(16):
                    kx = i - j
                    if 0 <= kx < 7:
                        kbase = ky * 64 + (i % 8) * 8 + (i // 8) * 4 + cin_off
                        wm[j * 32 + n_off:j * 32 + n_off + co, kbase:kbase + ci] = w[:, :, ky, kx].float()
    return wm


class StemSuperSpec:
    """The dual stem as ONE gather-GEMM over super-pixels: M = pixels / 4, N = 128, K = 7 blocks of 64 -- the same
    FLOPs as the per-pixel form (N = 32) but a quarter of the rows, so a quarter of the producers' LDG -> STS traffic
    and MMAs that are 4x wider per A byte.  Epilogue EPI_STEM_SP writes the 64-channel PL super-pixel view that the
    level0 super-pixel convolution (`superpixel_weight`) reads by TMA."""

    def __init__(self, wi, wh, scale, shift, in_W, nplanes):
        dev = wi.device
        Wp = in_W + 6                                   # the stem input frame has a 3-pixel border (= the padding)
        seg_off = []
        for ky in range(7):
            seg_off += [ky * Wp, ky * Wp + 8]           # two 8-pixel (64-byte) segments of kernel row ky
        self.seg_off = (ctypes.c_int * 14)(*seg_off)
        self.nkb, self.seg_groups = 7, 4
        wm = stem_superpixel_matrix(wi.float(), wh.float()).to(dev)
        self.scale = scale.float().contiguous()
        self.shift = shift.float().contiguous()
        nbytes = _lib.load().sgta_planes_wpack_bytes(128, 7 * 64, nplanes)
        self.wpack = torch.empty(nbytes, device=dev, dtype=torch.uint8)
        _lib.call("sgta_planes_pack_weight", _lib.ptr(wm.contiguous()), _lib.ptr(self.wpack), 128, 7 * 64, nplanes, _lib.stream())
        self.Cout = 128


def conv_stem_sp(spec, x, y):
    """x: SC view, 4 channels, border 3, [B,H,W];  y: PL view, 64 channels, [B,H,W/4]."""
    if x.W % SP_G:
        raise _lib.SgtaError("super-pixel stem: W must be a multiple of 4")
    _lib.call("sgta_planes_conv_sc", x.ref, _lib.ptr(spec.wpack), _lib.ptr(spec.scale), _lib.ptr(spec.shift), y.ref,
              spec.Cout, 1, SP_G, x.H, x.W // SP_G, spec.nkb, spec.seg_groups, spec.seg_off, ACT_NONE, EPI_STEM_SP,
              _lib.stream())


def dcn(x, om, wspec, scale, shift, y, relu=True):
    """x, y: PlaneViews (PL); om: fp32 [rows >= P, 32] raw conv_offset_mask output."""
    _lib.call("sgta_planes_dcn", x.ref, _lib.ptr(om), _lib.ptr(wspec.wpack), _lib.ptr(scale), _lib.ptr(shift), y.ref,
              wspec.Cin, wspec.Cout, int(relu), _lib.stream())


def pack_stem(img, hm, y, b_off):
    _lib.call("sgta_planes_pack_stem", _lib.ptr(img), _lib.ptr(hm), y.ref, b_off, img.shape[0], _lib.stream())


def maxpool2(x, y, C, xc_off=0, yc_off=0):
    _lib.call("sgta_planes_maxpool2", x.ref, xc_off, y.ref, yc_off, C, _lib.stream())


def upsample_add(x, w_up, skip, y, C, f):
    _lib.call("sgta_planes_upsample_add", x.ref, _lib.ptr(w_up), skip.ref if skip is not None else None, y.ref, C, f,
              _lib.stream())


def superpixels(sc, sp, to_super=True):
    """SC view (16 channels, [B,H,W]) <-> PL super-pixel view (64 channels, [B,H,W/4]); see superpixel_weight."""
    _lib.call("sgta_planes_superpixels", sc.ref, sp.ref, int(bool(to_super)), _lib.stream())


def gather_tokens(x, b_off, ids, C):
    B, n = ids.shape
    rows = torch.empty(B, n, C, device=ids.device, dtype=torch.float32)
    _lib.call("sgta_planes_gather_tokens", x.ref, b_off, _lib.ptr(ids), _lib.ptr(rows), B, C, n, _lib.stream())
    return rows


def scatter_tokens(x, b_off, ids, rows, C):
    B, n = ids.shape
    rows = rows.contiguous().float()
    _lib.call("sgta_planes_scatter_tokens", x.ref, b_off, _lib.ptr(ids), _lib.ptr(rows), B, C, n, _lib.stream())
