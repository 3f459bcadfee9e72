"""Image pre-processing on the device for a lock-step batch of clips.

Mirrors `SGTADetector._transform_scale` / `pre_process` / `normalize_img`
(sgtapose/lib/sgta_detector.py:334-366, :368-399, :402-403) for test scale 1 in all three testing modes
(`fix_res`, the one the reference runs in, opts_parallel.py:341; `fix_short`; keep-resolution): same `meta` keys, same matrices (the 3-point solve stays
`cv2.getAffineTransform` on the host through `get_affine_transform`, lib/utils/image.py:45-78), and
the network input computed by ONE launch of `sgta_preprocess` from the raw uint8 frames -- cv2's
fixed-point bilinear warp restated bit-exactly, then ((img / 255.) - mean) / std in float32.

The reference uploads a float32 [1,3,H,W] tensor per frame (:154); here the raw frames travel as
uint8 (h*w*3 bytes instead of H*W*12).  No CPU fallback: inputs must reach the device.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from . import priors as PR

MEAN = (0.5, 0.5, 0.5)      # sgta_detector.py:58
STD = (0.5, 0.5, 0.5)       # sgta_detector.py:59


# lib/utils/image.py:45-78 and utilities.py:889-925 are the same function in the reference: one host restatement
get_affine_transform = PR.get_affine_transform


def transform_meta(height, width, opt, scale=1):
    """_transform_scale (sgta_detector.py:334-366, all three testing modes) + the meta dict of pre_process
    (:375-394) for one raw frame size.  `fix_res` (the mode the reference runs in, opts_parallel.py:341) is the
    default; `fix_short` and keep-resolution (`fix_res` False, padded to `opt.pad`) follow the same lines."""
    new_height, new_width = int(height * scale), int(width * scale)
    if (new_height, new_width) != (height, width):
        # cv2.resize before the warp (:365) is a second interpolation the device kernel does not restate
        raise _lib.SgtaError("pre_process: only test scale 1 is built (the reference's default test_scales)")
    fix_short = int(getattr(opt, "fix_short", -1))
    if fix_short > 0:
        if height < width:
            inp_h, inp_w = fix_short, (int(width / height * fix_short) + 63) // 64 * 64
        else:
            inp_h, inp_w = (int(height / width * fix_short) + 63) // 64 * 64, fix_short
        c = np.array([width / 2, height / 2], dtype=np.float32)
        s = np.array([width, height], dtype=np.float32)
    elif getattr(opt, "fix_res", True):
        inp_h, inp_w = int(opt.input_h), int(opt.input_w)
        c = np.array([new_width / 2., new_height / 2.], dtype=np.float32)
        s = max(height, width) * 1.0
    else:
        pad = int(getattr(opt, "pad", 31))
        inp_h, inp_w = (new_height | pad) + 1, (new_width | pad) + 1
        c = np.array([new_width // 2, new_height // 2], dtype=np.float32)
        s = np.array([inp_w, inp_h], dtype=np.float32)
    down = int(getattr(opt, "down_ratio", 4))
    out_h, out_w = inp_h // down, inp_w // down
    return {"c": c, "s": s, "height": height, "width": width, "out_height": out_h, "out_width": out_w,
            "inp_height": inp_h, "inp_width": inp_w,
            "trans_input": get_affine_transform(c, s, 0, [inp_w, inp_h]),
            "trans_output": get_affine_transform(c, s, 0, [out_w, out_h])}


def warp_normalize(frames_u8, trans, out_hw, mean=MEAN, std=STD, return_u8=False, out=None):
    """frames_u8 [B,h,w,3] uint8 CUDA, trans [6] / [2,3] (shared) or [B,2,3] forward matrices
    -> network input [B,3,H,W] float32 (and the warped uint8 image [B,H,W,3] if asked).
    `out`: an existing contiguous [B,3,H,W] float32 CUDA tensor to write into (the engine's input buffer)."""
    if not frames_u8.is_cuda:
        raise _lib.SgtaError("pre_process: frames must be CUDA tensors (no CPU fallback)")
    if frames_u8.dtype != torch.uint8 or frames_u8.dim() != 4 or frames_u8.shape[3] != 3:
        raise _lib.SgtaError("pre_process: frames must be uint8 [B,h,w,3]")
    frames_u8 = frames_u8.contiguous()
    B, h, w, _ = frames_u8.shape
    H, W = int(out_hw[0]), int(out_hw[1])
    t = np.ascontiguousarray(np.asarray(trans, np.float64).reshape(-1, 6))
    if out is None:
        out = torch.empty(B, 3, H, W, device=frames_u8.device, dtype=torch.float32)
    elif tuple(out.shape) != (B, 3, H, W) or out.dtype != torch.float32 or not out.is_contiguous():
        raise _lib.SgtaError("pre_process: `out` must be a contiguous float32 [B,3,H,W] tensor")
    u8 = torch.empty(B, H, W, 3, device=frames_u8.device, dtype=torch.uint8) if return_u8 else None
    m3 = (ctypes.c_float * 3)(*[float(v) for v in mean])
    s3 = (ctypes.c_float * 3)(*[float(v) for v in std])
    _lib.call("sgta_preprocess", _lib.ptr(frames_u8), _lib.ptr(out), _lib.ptr(u8),
              t.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), t.shape[0], m3, s3, B, h, w, H, W, _lib.stream())
    return (out, u8) if return_u8 else out


def pre_process(frames, opt, device="cuda"):
    """SGTADetector.pre_process for B frames of one raw size.
    frames: uint8 [B,h,w,3] (numpy or torch, host or device; host frames are uploaded as uint8).
    -> (images [B,3,H,W] float32 on the device, meta)."""
    if isinstance(frames, np.ndarray):
        frames = torch.from_numpy(np.ascontiguousarray(frames))
    if frames.dim() == 3:
        frames = frames.unsqueeze(0)
    if not frames.is_cuda:
        frames = frames.to(device, non_blocking=True)
    meta = transform_meta(int(frames.shape[1]), int(frames.shape[2]), opt)
    images = warp_normalize(frames, meta["trans_input"], (meta["inp_height"], meta["inp_width"]))
    return images, meta
