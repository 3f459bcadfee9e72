"""Heatmap decode on the device, behind the reference's function names.

  dream_generic_decode(output, K, opt)   reference sgtapose/lib/model/decode.py:184-313
  _peaks_info(scores)                    lib/model/utils.py:207-284 (+ image_proc.py:1032-1143)
  generic_decode / _nms / _topk          decode.py:83-182, utils.py:59-103
  SoftArgmaxPavlo                        sgtapose/spatial_softmax.py:15-95

Same dict keys and shapes as the reference.  Differences, all deliberate:
  * every sample of the batch is decoded (the reference reads scores[0] only, utils.py:214);
  * nothing leaves the device: no scipy, no .cpu() per keypoint;
  * 'scores' is returned on the device (the reference builds it on the CPU, utils.py:271).
Integer outputs (inds, xs, ys) are bit-exact w.r.t. the reference arithmetic on identical
heatmaps: the kernel reproduces scipy's float64 gaussian_filter and numpy.average.
"""
import ctypes

import numpy as np
import torch

from . import _lib

_SIGMA, _RADIUS = 3, 12


def _gauss_weights():
    # scipy.ndimage._filters._gaussian_kernel1d(sigma=3, order=0, radius=int(4*3+0.5)), float64
    x = np.arange(-_RADIUS, _RADIUS + 1)
    phi = np.exp(-0.5 / (_SIGMA * _SIGMA) * x ** 2)
    w = np.ascontiguousarray(phi / phi.sum(), dtype=np.float64)
    return w


_GW = _gauss_weights()
_GW_PTR = _GW.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _prep(t):
    if not t.is_cuda:
        raise _lib.SgtaError("decode: inputs must be CUDA tensors (no CPU fallback)")
    return t.contiguous().float()


def recheck_count(reset=True):
    """Pixels the production decode re-evaluated in float64 since the last reset (test hook)."""
    n = ctypes.c_uint64(0)
    _lib.call("sgta_decode_recheck_count", ctypes.byref(n), int(reset))
    return int(n.value)


def full_map(on):
    """Test hook: make the production decode blur every pixel instead of the active box only; returns the old setting."""
    return bool(_lib.load().sgta_decode_full_map(int(bool(on))))


def peaks_decode(hm, reg=None, tracking=None, exact64=False):
    """One launch of the live decode.  Returns dict of device tensors:
    scores [B,C] f32, inds/xs/ys [B,C] i64, cts_wreg [B,C,2] f32, tracking [B,C,2] f32|None.
    exact64=True runs the all-float64 cross-check kernel instead of the production one."""
    hm = _prep(hm)
    B, C, h, w = hm.shape
    reg = _prep(reg) if reg is not None else None
    tracking = _prep(tracking) if tracking is not None else None
    dev = hm.device
    scores = torch.empty(B, C, device=dev, dtype=torch.float32)
    inds = torch.empty(B, C, device=dev, dtype=torch.int64)
    xs, ys = torch.empty_like(inds), torch.empty_like(inds)
    cts_wreg = torch.empty(B, C, 2, device=dev, dtype=torch.float32)
    trk = torch.empty(B, C, 2, device=dev, dtype=torch.float32) if tracking is not None else None
    _lib.call("sgta_decode_peaks_exact64" if exact64 else "sgta_decode_peaks", _lib.ptr(hm), _lib.ptr(reg), _lib.ptr(tracking), _lib.ptr(scores),
              _lib.ptr(inds), _lib.ptr(xs), _lib.ptr(ys), _lib.ptr(cts_wreg), _lib.ptr(trk), _GW_PTR,
              B, C, h, w, _lib.stream())
    return {"scores": scores, "inds": inds, "xs": xs, "ys": ys, "cts_wreg": cts_wreg, "tracking": trk}


def _peaks_info(scores):
    """utils.py:207-284 -> (topk_score, topk_inds, topk_clses, topk_ys, topk_xs)."""
    B, C = scores.shape[:2]
    r = peaks_decode(scores)
    clses = torch.arange(C, device=scores.device).view(1, C).expand(B, C).contiguous()
    return r["scores"], r["inds"], clses, r["ys"], r["xs"]


def dream_generic_decode(output, K=7, opt=None):
    """decode.py:184-313 for the heads this model has (hm, reg, tracking)."""
    if "hm" not in output:
        return {}
    if opt is not None and getattr(opt, "zero_tracking", False):
        output["tracking"] *= 0
    heat = output["hm"]
    B, C, h, w = heat.shape
    r = peaks_decode(heat, output.get("reg"), output.get("tracking"))
    xs0, ys0 = r["xs"], r["ys"]
    clses = torch.arange(C, device=heat.device).view(1, C).expand(B, C)
    cts = torch.stack([xs0, ys0], dim=2)
    ret = {"scores": r["scores"].view(B, K), "clses": clses.reshape(B, K).float(), "xs": xs0, "ys": ys0,
           "cts": cts, "inds": r["inds"]}
    cts_wreg = r["cts_wreg"].view(B, K, 2, 1)          # the reference's cat(dim=2) shape, decode.py:232
    ret["cts_wreg"] = cts_wreg
    ret["regs"] = cts_wreg - cts.view(B, K, 2, 1).float()
    if r["tracking"] is not None:
        ret["tracking"] = r["tracking"].view(B, K, -1)
    return ret


def _nms(heat, kernel=3):
    """utils.py:59-65."""
    if kernel != 3:
        raise _lib.SgtaError("_nms: only the 3x3 window the reference uses is built")
    heat = _prep(heat)
    B, C, h, w = heat.shape
    out = torch.empty_like(heat)
    _lib.call("sgta_nms3x3", _lib.ptr(heat), _lib.ptr(out), B, C, h, w, _lib.stream())
    return out


def nms_topk(heat, K):
    """_nms followed by _topk in one pass over the heatmap.
    -> (scores [B,K], inds [B,K] i64, clses [B,K] i32, ys [B,K] f32, xs [B,K] f32)"""
    heat = _prep(heat)
    B, C, h, w = heat.shape
    dev = heat.device
    scores = torch.empty(B, K, device=dev, dtype=torch.float32)
    inds = torch.empty(B, K, device=dev, dtype=torch.int64)
    clses = torch.empty(B, K, device=dev, dtype=torch.int32)
    ws = torch.empty(B * C * K * 3, device=dev, dtype=torch.float32)
    _lib.call("sgta_decode_nms_topk", _lib.ptr(heat), _lib.ptr(scores), _lib.ptr(inds), _lib.ptr(clses),
              _lib.ptr(ws), B, C, h, w, K, _lib.stream())
    ys = torch.div(inds, w, rounding_mode="floor").float()
    xs = (inds % w).float()
    return scores, inds, clses, ys, xs


def generic_decode(output, K=7, opt=None):
    """decode.py:83-182 for the heads this model has."""
    if "hm" not in output:
        return {}
    if opt is not None and getattr(opt, "zero_tracking", False):
        output["tracking"] *= 0
    heat = output["hm"]
    B, C, h, w = heat.shape
    scores, inds, clses, ys0, xs0 = nms_topk(heat, K)
    ret = {"scores": scores, "clses": clses.float(), "xs": xs0, "ys": ys0,
           "cts": torch.stack([xs0, ys0], dim=2), "inds": inds}
    for head in ("reg", "tracking"):
        if head in output:
            f = output[head]
            flat = f.reshape(B, f.shape[1], h * w)
            ret[head] = torch.gather(flat, 2, inds[:, None, :].expand(-1, f.shape[1], -1)).permute(0, 2, 1)
    return ret


class SoftArgmaxPavlo(torch.nn.Module):
    """spatial_softmax.py:15-95 (fixed beta; learned_beta is outside the hot path)."""

    def __init__(self, n_keypoints=5, learned_beta=False, initial_beta=25.0):
        super().__init__()
        if learned_beta:
            raise _lib.SgtaError("SoftArgmaxPavlo: learned_beta is not built")
        self.beta = float(initial_beta)

    def forward(self, heatmaps, size_mult=1.0):
        hm = _prep(heatmaps)
        B, C, h, w = hm.shape
        out = torch.empty(B, C, 2, device=hm.device, dtype=torch.float32)
        _lib.call("sgta_soft_argmax", _lib.ptr(hm), _lib.ptr(out), B, C, h, w, self.beta,
                  float(size_mult), _lib.stream())
        return out
