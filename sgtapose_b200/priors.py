"""Structure-prior maps on the device.

Host side of csrc/priors.cu, mirroring the reference's helpers by name:
  get_affine_transform          sgtapose/utilities.py:889-925 (cv2.getAffineTransform of 3 point pairs)
  affine_transform_and_clip     sgtapose/utilities.py:943-972 (float64, host: 7 points per clip)
  get_prev_hm_wo_noise[_cls]    sgtapose/utilities.py:1045-1057, :1085-1098 -> rendered on the DEVICE
The reference renders four numpy maps per clip and frame and uploads them (lib/sgta_detector.py:528-540,
1.7 MB per clip at 384^2); here the host ships 2 x 7 float64 centres per clip and one kernel launch writes
the maps of the whole lock-step batch straight into the engine's input buffers.
"""
import ctypes

import numpy as np
import torch

from . import _lib

RADIUS, SIGMA = 4, 2.0


def _gaussian2d():
    # utilities.py:846-853 with shape (9, 9), sigma 2, res (0, 0); float64 then rounded to fp32 like
    # np.maximum(float32 map, float64 gaussian, out=float32 map) does
    m = float(RADIUS)
    y, x = np.ogrid[-m:m + 1, -m:m + 1]
    h = np.exp(-(x * x + y * y) / (2 * SIGMA * SIGMA))
    h[h < np.finfo(h.dtype).eps * h.max()] = 0
    return np.ascontiguousarray(h, dtype=np.float32)


_G = _gaussian2d()
_G_PTR = _G.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def get_affine_transform(center, scale, rot, output_size, shift=np.array([0, 0], dtype=np.float32), inv=0):
    """utilities.py:889-925, same arguments."""
    import cv2
    if not isinstance(scale, (np.ndarray, list)):
        scale = np.array([scale, scale], dtype=np.float32)
    src_w, dst_w, dst_h = scale[0], output_size[0], output_size[1]
    rot_rad = np.pi * rot / 180
    sn, cs = np.sin(rot_rad), np.cos(rot_rad)
    p = [0, src_w * -0.5]
    src_dir = np.array([p[0] * cs - p[1] * sn, p[0] * sn + p[1] * cs])          # get_dir, utilities.py:927-935
    dst_dir = np.array([0, dst_w * -0.5], np.float32)
    src = np.zeros((3, 2), dtype=np.float32)
    dst = np.zeros((3, 2), dtype=np.float32)
    src[0, :] = center + scale * shift
    src[1, :] = center + src_dir + scale * shift
    dst[0, :] = [dst_w * 0.5, dst_h * 0.5]
    dst[1, :] = np.array([dst_w * 0.5, dst_h * 0.5], np.float32) + dst_dir
    for a in (src, dst):
        d = a[0, :] - a[1, :]
        a[2:, :] = a[1, :] + np.array([-d[1], d[0]], dtype=np.float32)
    if inv:
        return cv2.getAffineTransform(np.float32(dst), np.float32(src))
    return cv2.getAffineTransform(np.float32(src), np.float32(dst))


def affine_transform_and_clip(pts, t, width, height, raw_width, raw_height):
    """utilities.py:943-972 for pts [..., n_kp, 2] (vectorised over leading dims; same float64 arithmetic:
    np.dot of the 2x3 matrix with [x, y, 1] per point)."""
    pts = np.asarray(pts, np.float64)
    lead = pts.shape[:-1]
    flat = pts.reshape(-1, 2)
    hom = np.concatenate((flat, np.ones((flat.shape[0], 1))), axis=-1)
    new = np.dot(t, hom.T).T
    new[:, 0] = np.clip(new[:, 0], 0, width - 1)
    new[:, 1] = np.clip(new[:, 1], 0, height - 1)
    inside = (flat[:, 0] >= 0.0) & (flat[:, 0] < raw_width) & (flat[:, 1] >= 0.0) & (flat[:, 1] < raw_height)
    new[~inside] = 0
    return new.reshape(lead + (2,))


def render_priors(centres_in, centres_out, size, out_size, hm=None, hm_cls=None):
    """centres_*: [B,K,2] float64 (numpy or torch, host or device) in input / output pixels, or None to
    skip that map.  hm [B,1,size,size] / hm_cls [B,K,out_size,out_size]: fp32 CUDA tensors to fill
    (allocated if omitted).  Returns (hm, hm_cls)."""
    ref = centres_in if centres_in is not None else centres_out
    if ref is None:
        raise _lib.SgtaError("render_priors: no centres given")
    B, K = int(ref.shape[0]), int(ref.shape[1])
    dev = hm.device if hm is not None else hm_cls.device if hm_cls is not None else torch.device("cuda")

    def dev64(c):
        if c is None:
            return None
        t = torch.as_tensor(np.ascontiguousarray(c) if isinstance(c, np.ndarray) else c, dtype=torch.float64)
        return t.to(dev, non_blocking=True).contiguous()
    ci, co = dev64(centres_in), dev64(centres_out)
    if ci is not None and hm is None:
        hm = torch.empty(B, 1, size, size, device=dev, dtype=torch.float32)
    if co is not None and hm_cls is None:
        hm_cls = torch.empty(B, K, out_size, out_size, device=dev, dtype=torch.float32)
    _lib.call("sgta_render_priors", _lib.ptr(ci), _lib.ptr(co), _lib.ptr(hm) if ci is not None else None,
              _lib.ptr(hm_cls) if co is not None else None, _G_PTR, B, K, size, size, out_size, out_size,
              _lib.stream())
    return hm, hm_cls


def get_prev_hm_wo_noise(kp_projs_raw, trans_input, input_w, input_h, raw_width, raw_height, device="cuda"):
    """utilities.py:1045-1057 for ONE clip, on the device: -> [input_h, input_w] fp32 CUDA tensor."""
    if input_w != input_h:
        raise _lib.SgtaError("square maps only (dla.py:931 asserts H == W)")
    if kp_projs_raw is None:
        return torch.zeros(input_h, input_w, device=device)
    c = affine_transform_and_clip(kp_projs_raw, trans_input, input_w, input_h, raw_width, raw_height)
    hm = torch.empty(1, 1, input_h, input_w, device=device)
    render_priors(c[None], None, input_w, input_w, hm=hm)
    return hm[0, 0]


def get_prev_hm_wo_noise_cls(kp_projs_raw, kp_gts_raw, trans_input, input_w, input_h, raw_width, raw_height,
                             device="cuda"):
    """utilities.py:1085-1098 for ONE clip, on the device: -> [n_kp, input_h, input_w] fp32 CUDA tensor."""
    n_kp = int(np.asarray(kp_gts_raw).shape[0])
    if kp_projs_raw is None:
        return torch.zeros(n_kp, int(input_h), int(input_w), device=device)
    c = affine_transform_and_clip(kp_projs_raw, trans_input, input_w, input_h, raw_width, raw_height)
    cls = torch.empty(1, n_kp, int(input_h), int(input_w), device=device)
    render_priors(None, c[None], int(input_w), int(input_w), hm_cls=cls)
    return cls[0]
