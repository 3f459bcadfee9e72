"""Deterministic synthetic weights and inputs.

No dataset or checkpoint exists offline, so benchmarks and parity tests use
  * a synthetic state-dict: every tensor filled from a CPU torch.Generator seeded by a hash
    of (seed, key name) -- independent of module construction order, so the reference model,
    the oracle and this package all load the identical weights; zero-initialised parameters
    (conv_offset_mask.*, pos_embed) and BN running stats are randomised so every path is
    exercised;
  * synthetic Panda-like inputs: N(0,1) images and prior heatmaps rendered like
    utilities.py:800-853 (radius 4, sigma 2, integer centre, max-blend).
"""
import hashlib
import math

import numpy as np
import torch

N_KP = 7


def _gen(seed, key):
    h = hashlib.sha256(("%d:%s" % (seed, key)).encode()).digest()
    g = torch.Generator(device="cpu")
    g.manual_seed(int.from_bytes(h[:7], "little"))
    return g


def synthetic_state_dict(template, seed=317):
    """template: a state_dict (only keys/shapes/dtypes are read).  Returns a new CPU dict."""
    out = {}
    shared = {}
    for k, t in template.items():
        # transformer.X.layers.{0,1,2}.* are three aliases of one tensor (dla.py:788-789)
        alias = k
        if ".layers." in k:
            head, tail = k.split(".layers.")
            alias = head + ".layers.0." + tail.split(".", 1)[1]
        if alias in shared:
            out[k] = shared[alias]
            continue
        g = _gen(seed, alias)
        shape = tuple(t.shape)
        if k.endswith("num_batches_tracked"):
            v = torch.zeros(shape, dtype=t.dtype)
        elif k.endswith("running_mean"):
            v = torch.randn(shape, generator=g) * 0.1
        elif k.endswith("running_var"):
            v = torch.rand(shape, generator=g) + 0.5
        elif k.endswith("pos_embed"):
            v = torch.randn(shape, generator=g) * 0.02
        elif "conv_offset_mask.weight" in k:
            v = torch.randn(shape, generator=g) * 0.01
        elif "conv_offset_mask.bias" in k:
            v = torch.rand(shape, generator=g) - 0.5
        elif k == "hm.2.bias":
            v = torch.full(shape, -4.6)
        elif len(shape) == 1 and (".bn" in k or "actf" in k or ".norm" in k or k.split(".")[-2].isdigit()) \
                and k.endswith("weight"):
            v = torch.rand(shape, generator=g) * 0.5 + 0.75          # BN / LN gains
        elif len(shape) == 1:
            v = (torch.rand(shape, generator=g) - 0.5) * 0.2          # biases
        elif ".up_" in k:
            v = None                                                  # keep bilinear init below
        else:
            fan_in = int(np.prod(shape[1:]))
            bound = math.sqrt(3.0 / fan_in) * 1.4
            v = (torch.rand(shape, generator=g) * 2 - 1) * bound
        if v is None:
            kk = shape[2]
            f = math.ceil(kk / 2)
            c = (2 * f - 1 - f % 2) / (2.0 * f)
            ramp = torch.tensor([1 - abs(i / f - c) for i in range(kk)])
            v = (ramp[:, None] * ramp[None, :])[None, None].expand(shape).clone()
            v = v * (1 + 0.1 * (torch.rand(shape[0], 1, 1, 1, generator=g) - 0.5))
        v = v.to(t.dtype).contiguous()
        shared[alias] = v
        out[k] = v
    return out


def draw_gaussian(hm, cx, cy, radius=4, sigma=2.0):
    """utilities.py:800-824 draw_umich_gaussian: integer centre, drawn only if the whole
    (2r+1)^2 patch (+1) fits, max-blend."""
    h, w = hm.shape
    x, y = int(cx), int(cy)
    if not (x - radius >= 0 and x + radius + 1 < w and y - radius >= 0 and y + radius + 1 < h):
        return hm
    r = np.arange(-radius, radius + 1, dtype=np.float64)
    g = np.exp(-(r[None] ** 2 + r[:, None] ** 2) / (2 * sigma * sigma))
    g[g < np.finfo(g.dtype).eps * g.max()] = 0
    patch = hm[y - radius:y + radius + 1, x - radius:x + radius + 1]
    np.maximum(patch, g.astype(hm.dtype), out=patch)
    return hm


def synthetic_keypoints(B, S, seed=317, frame=0):
    """Smooth pseudo-trajectory of 7 keypoints in input-pixel coordinates, [B,7,2]."""
    rng = np.random.default_rng(seed)
    base = rng.uniform(0.2 * S, 0.8 * S, size=(B, N_KP, 2))
    vel = rng.uniform(-3, 3, size=(B, N_KP, 2))
    return base + vel * frame


def synthetic_inputs(B, S=384, seed=317, frame=1, dtype=torch.float32):
    """x, pre_img [B,3,S,S]; pre_hm, repro_hm [B,1,S,S]; pre_hm_cls, repro_hm_cls [B,7,S/4,S/4]."""
    g = _gen(seed, "inputs:%d:%d:%d" % (B, S, frame))
    x = torch.randn(B, 3, S, S, generator=g)
    pre_img = torch.randn(B, 3, S, S, generator=g)
    kp_prev = synthetic_keypoints(B, S, seed, frame - 1)
    kp_cur = synthetic_keypoints(B, S, seed, frame)
    q = S // 4
    pre_hm = np.zeros((B, 1, S, S), np.float32)
    repro_hm = np.zeros((B, 1, S, S), np.float32)
    pre_cls = np.zeros((B, N_KP, q, q), np.float32)
    repro_cls = np.zeros((B, N_KP, q, q), np.float32)
    for b in range(B):
        for k in range(N_KP):
            draw_gaussian(pre_hm[b, 0], *kp_prev[b, k])
            draw_gaussian(repro_hm[b, 0], *kp_cur[b, k])
            draw_gaussian(pre_cls[b, k], *(kp_prev[b, k] / 4))
            draw_gaussian(repro_cls[b, k], *(kp_cur[b, k] / 4))
    t = lambda a: torch.from_numpy(a).to(dtype)
    return (x.to(dtype), pre_img.to(dtype), t(pre_hm), t(repro_hm), t(pre_cls), t(repro_cls))


def synthetic_heatmaps(B, C=7, h=96, w=96, seed=317, noise=0.005, missing_every=0):
    """Decode-only inputs: sigma-2 Gaussian blobs at random centres + U(0,noise)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    hm = rng.random((B, C, h, w), dtype=np.float32) * np.float32(noise)
    centres = rng.uniform(3, min(h, w) - 4, size=(B, C, 2)).astype(np.float32)
    for b in range(B):
        for c in range(C):
            if missing_every and (b * C + c) % missing_every == missing_every - 1:
                continue
            cx, cy = centres[b, c]
            hm[b, c] += np.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / 8.0).astype(np.float32)
    return torch.from_numpy(np.clip(hm, 0, 1)), centres
