"""Oracle: heatmap decode.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  numpy restatements of

* the LIVE decode ``dream_generic_decode`` -> ``_peaks_info`` ->
  ``peaks_from_belief_maps`` (reference ``sgtapose/lib/model/decode.py:184-313``,
  ``sgtapose/lib/model/utils.py:207-284``, ``sgtapose/image_proc.py:1032-1143``),
  including a bit-exact restatement of ``scipy.ndimage.gaussian_filter(sigma=3)``
  on float32 maps and of ``numpy.average``'s summation order;
* the alternate decode named by north_star: ``_nms`` + ``_topk``
  (``utils.py:59-103``, used by ``generic_decode`` ``decode.py:83-182``);
* ``SoftArgmaxPavlo`` (``sgtapose/spatial_softmax.py:15-95``).

Pinned by tests/test_oracle_decode.py against scipy/numpy themselves and against
golden vectors generated from the imported reference (tests/golden/decode_*.npz).

Batch generalisation: the reference ``_peaks_info`` reads ``scores[0]`` only
(``utils.py:214``); here every sample of the batch is decoded by the same rule.
"""
import numpy as np

SIGMA = 3
TRUNCATE = 4.0
RADIUS = int(TRUNCATE * SIGMA + 0.5)          # 12  (scipy _gaussian_kernel1d radius)
THRESH_BLURRED = 0.01                         # image_proc.py:1043
OFFSET_DUE_TO_UPSAMPLING = 0.4395             # utils.py:212
AMBIGUITY_GAP = 0.25                          # utils.py:230-233
WIN = 5                                       # image_proc.py:1079


def gaussian_weights(sigma=SIGMA, radius=RADIUS):
    """scipy.ndimage._filters._gaussian_kernel1d(sigma, 0, radius) in float64."""
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / (sigma * sigma) * x ** 2)
    return phi / phi.sum()


def _reflect(i, n):
    """scipy 'reflect' (half-sample symmetric): -1->0, -2->1, n->n-1, n+1->n-2 ..."""
    if n == 1:
        return 0
    period = 2 * n
    i = i % period
    if i < 0:
        i += period
    return i if i < n else period - 1 - i


def _correlate1d_symmetric(a, w, axis):
    """scipy ``correlate1d`` symmetric-kernel branch (ni_filters.c NI_Correlate1D):
    t = x[l]*w[r]; for ii in -r..-1: t += (x[l+ii] + x[l-ii]) * w[ii+r]; float64
    accumulate, output cast to the input dtype (float32)."""
    r = (len(w) - 1) // 2
    a = np.moveaxis(a, axis, -1)
    n = a.shape[-1]
    idx = np.array([_reflect(i, n) for i in range(-r, n + r)], dtype=np.int64)
    ext = a[..., idx].astype(np.float64)                      # [..., n + 2r]
    t = ext[..., r:r + n] * w[r]
    for ii in range(-r, 0):
        t = t + (ext[..., r + ii:r + ii + n] + ext[..., r - ii:r - ii + n]) * w[ii + r]
    return np.moveaxis(t.astype(a.dtype), -1, axis)


def gaussian_blur(map32):
    """scipy.ndimage.gaussian_filter(map32, sigma=3) for a float32 [h, w] map
    (image_proc.py:1053): separable, axis 0 first, float32 intermediate."""
    w = gaussian_weights()
    out = _correlate1d_symmetric(np.asarray(map32, dtype=np.float32), w, 0)
    return _correlate1d_symmetric(out, w, 1)


def _np_sum_1d(v):
    """numpy pairwise summation order for a contiguous 1-D float64 array, n < 128
    (numpy/core/src/umath/loops_utils.h.src pairwise_sum): 8 interleaved partial
    sums over the first n - n%8 elements, tree-combined, then the tail in order."""
    n = len(v)
    if n < 8:
        res = 0.0
        for x in v:
            res = res + x
        return res
    r = [v[j] for j in range(8)]
    i = 8
    while i < n - (n % 8):
        for j in range(8):
            r[j] = r[j] + v[i + j]
        i += 8
    res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]))
    while i < n:
        res = res + v[i]
        i += 1
    return res


def _centroid(map_ori, px, py):
    """image_proc.py:1082-1116: intensity-weighted centroid over the 5x5 window of
    the UN-blurred map (np.average with weights; out-of-map cells have weight 0 and
    coordinate 0), + 0.4395; falls back to the integer peak if all weights are 0."""
    h, w = map_ori.shape
    ran = WIN // 2
    weights = np.zeros(WIN * WIN, dtype=np.float64)
    iv = np.zeros(WIN * WIN, dtype=np.float64)
    jv = np.zeros(WIN * WIN, dtype=np.float64)
    for i in range(-ran, ran + 1):
        for j in range(-ran, ran + 1):
            if py + i < 0 or py + i >= h or px + j < 0 or px + j >= w:
                continue
            flat = (j + ran) * WIN + (i + ran)       # arrays indexed [j+ran, i+ran]
            iv[flat] = py + i
            jv[flat] = px + j
            weights[flat] = np.float64(map_ori[py + i, px + j])
    scl = _np_sum_1d(weights)
    if scl == 0.0:
        return px + OFFSET_DUE_TO_UPSAMPLING, py + OFFSET_DUE_TO_UPSAMPLING
    cx = _np_sum_1d(jv * weights) / scl + OFFSET_DUE_TO_UPSAMPLING
    cy = _np_sum_1d(iv * weights) / scl + OFFSET_DUE_TO_UPSAMPLING
    return cx, cy


def find_peaks(map_ori):
    """peaks_from_belief_maps for one channel -> list of (cx, cy, score) in
    np.nonzero (row-major) order.  Peak test on the blurred map, zero outside."""
    map_ori = np.asarray(map_ori, dtype=np.float32)
    m = gaussian_blur(map_ori)
    h, w = m.shape
    left = np.zeros((h, w)); left[1:, :] = m[:-1, :]
    right = np.zeros((h, w)); right[:-1, :] = m[1:, :]
    up = np.zeros((h, w)); up[:, 1:] = m[:, :-1]
    down = np.zeros((h, w)); down[:, :-1] = m[:, 1:]
    binary = (m >= left) & (m >= right) & (m >= up) & (m >= down) & (m > THRESH_BLURRED)
    ys, xs = np.nonzero(binary)
    out = []
    for px, py in zip(xs.tolist(), ys.tolist()):
        cx, cy = _centroid(map_ori, px, py)
        out.append((cx, cy, map_ori[py, px]))
    return out


def select_peak(map_ori, peaks):
    """utils.py:221-275: one peak -> take it; several -> stable sort by cy descending,
    accept the first iff score0 - score1 >= 0.25; else / none -> missing.
    Returns (score, x_int, y_int); missing = (-1.0, 0, 0)."""
    if len(peaks) == 1:
        cx, cy = peaks[0][0], peaks[0][1]
    elif len(peaks) > 1:
        srt = sorted(peaks, key=lambda p: p[1], reverse=True)
        if srt[0][2] - srt[1][2] >= AMBIGUITY_GAP:
            cx, cy = srt[0][0], srt[0][1]
        else:
            return np.float32(-1.0), 0, 0
    else:
        return np.float32(-1.0), 0, 0
    x_int, y_int = int(cx), int(cy)
    return np.float32(map_ori[y_int, x_int]), x_int, y_int


def peaks_info(hm):
    """_peaks_info generalised over the batch: hm [B, C, h, w] float32 ->
    scores [B,C] f32, inds [B,C] i64, clses [B,C] i64, ys [B,C] i64, xs [B,C] i64."""
    hm = np.asarray(hm, dtype=np.float32)
    B, C, h, w = hm.shape
    scores = np.zeros((B, C), np.float32)
    xs = np.zeros((B, C), np.int64)
    ys = np.zeros((B, C), np.int64)
    for b in range(B):
        for c in range(C):
            s, x, y = select_peak(hm[b, c], find_peaks(hm[b, c]))
            scores[b, c], xs[b, c], ys[b, c] = s, x, y
    clses = np.tile(np.arange(C, dtype=np.int64)[None], (B, 1))
    return scores, ys * w + xs, clses, ys, xs


def _gather_hw(feat, inds):
    """_tranpose_and_gather_feat (utils.py:29-33): feat [B,F,h,w], inds [B,K] -> [B,K,F]."""
    B, Fc, h, w = feat.shape
    flat = feat.reshape(B, Fc, h * w)
    return np.stack([flat[b][:, inds[b]].T for b in range(B)], 0)


def dream_generic_decode(hm, reg=None, tracking=None):
    """decode.py:184-313 on numpy inputs (hm already sigmoid-ed, sgta_detector.py:854-862).
    Returns the same keys/shapes: scores [B,K], clses [B,K] float, xs/ys [B,K] int64,
    cts [B,K,2] int64, cts_wreg [B,K,2,1], regs [B,K,2,1], tracking [B,K,2]."""
    scores, inds, clses, ys0, xs0 = peaks_info(hm)
    B, K = scores.shape
    ret = {"scores": scores, "clses": clses.astype(np.float32), "xs": xs0, "ys": ys0,
           "inds": inds, "cts": np.stack([xs0, ys0], 2)}
    if reg is not None:
        r = _gather_hw(np.asarray(reg, np.float32), inds)
        xs = xs0.astype(np.float32)[..., None] + r[:, :, 0:1]
        ys = ys0.astype(np.float32)[..., None] + r[:, :, 1:2]
    else:
        xs = xs0.astype(np.float32)[..., None] + np.float32(0.5)
        ys = ys0.astype(np.float32)[..., None] + np.float32(0.5)
    cts_wreg = np.concatenate([xs[:, :, None], ys[:, :, None]], 2)     # [B,K,2,1] (decode.py:232)
    ret["cts_wreg"] = cts_wreg
    base = np.stack([xs0, ys0], 2).astype(np.float32)[:, :, :, None]
    ret["regs"] = cts_wreg - base
    if tracking is not None:
        ret["tracking"] = _gather_hw(np.asarray(tracking, np.float32), inds)
    return ret


# --------------------------------------------------------------------------------------
# alternate decode: 3x3 max-pool NMS + two-stage top-K   (utils.py:59-103, decode.py:93-94)
# --------------------------------------------------------------------------------------
def nms(heat, kernel=3):
    """_nms: keep = (maxpool(heat) == heat); implicit -inf padding."""
    heat = np.asarray(heat, np.float32)
    B, C, h, w = heat.shape
    pad = (kernel - 1) // 2
    p = np.full((B, C, h + 2 * pad, w + 2 * pad), -np.inf, np.float32)
    p[:, :, pad:pad + h, pad:pad + w] = heat
    hmax = np.full_like(heat, -np.inf)
    for dy in range(kernel):
        for dx in range(kernel):
            hmax = np.maximum(hmax, p[:, :, dy:dy + h, dx:dx + w])
    return heat * (hmax == heat).astype(np.float32)


def _topk_desc(v, K):
    """top-K of a 1-D array, value descending, index ascending among ties
    (the tie order this build defines; torch only pins K=1)."""
    order = np.lexsort((np.arange(len(v)), -v.astype(np.float64)))
    return order[:K]


def topk(scores, K):
    """_topk: per-channel top-K over h*w, then top-K over the C*K candidates."""
    scores = np.asarray(scores, np.float32)
    B, C, h, w = scores.shape
    out_s = np.zeros((B, K), np.float32)
    out_i = np.zeros((B, K), np.int64)
    out_c = np.zeros((B, K), np.int32)
    for b in range(B):
        cs, ci = [], []
        for c in range(C):
            idx = _topk_desc(scores[b, c].reshape(-1), K)
            cs.append(scores[b, c].reshape(-1)[idx]); ci.append(idx)
        cs = np.concatenate(cs); ci = np.concatenate(ci)
        sel = _topk_desc(cs, K)
        out_s[b], out_i[b], out_c[b] = cs[sel], ci[sel], sel // K
    ys = (out_i // w).astype(np.float32)
    xs = (out_i % w).astype(np.float32)
    return out_s, out_i, out_c, ys, xs


def generic_decode(hm, reg=None, tracking=None, K=7):
    heat = nms(hm)
    scores, inds, clses, ys0, xs0 = topk(heat, K)
    ret = {"scores": scores, "clses": clses.astype(np.float32), "xs": xs0, "ys": ys0,
           "inds": inds, "cts": np.stack([xs0, ys0], 2)}
    if reg is not None:
        r = _gather_hw(np.asarray(reg, np.float32), inds)
        ret["xs_reg"] = xs0[..., None] + r[:, :, 0:1]
        ret["ys_reg"] = ys0[..., None] + r[:, :, 1:2]
    if tracking is not None:
        ret["tracking"] = _gather_hw(np.asarray(tracking, np.float32), inds)
    return ret


# --------------------------------------------------------------------------------------
# SoftArgmaxPavlo   (spatial_softmax.py:15-95 == utils.py:107-187)
# --------------------------------------------------------------------------------------
def soft_argmax(heatmaps, beta=25.0, size_mult=1.0):
    """7x7 avg-pool (stride 1, pad 3, zeros counted in the divisor) -> subtract the
    per-map max -> exp(beta * .) -> normalise by (sum + 1e-8) -> E[x], E[y].  float32."""
    hm = np.asarray(heatmaps, np.float32)
    B, C, h, w = hm.shape
    p = np.zeros((B, C, h + 6, w + 6), np.float32)
    p[:, :, 3:3 + h, 3:3 + w] = hm
    acc = np.zeros_like(hm)
    for dy in range(7):
        for dx in range(7):
            acc = acc + p[:, :, dy:dy + h, dx:dx + w]
    pooled = acc / np.float32(49.0)
    flat = pooled.reshape(B, C, -1)
    flat = flat - flat.max(2, keepdims=True)
    e = np.exp(np.float32(beta) * flat)
    norm = e / (e.sum(2, keepdims=True) + np.float32(1e-8))
    norm = norm.reshape(B, C, h, w)
    col = (np.arange(w, dtype=np.float32) * np.float32(size_mult))[None, None, None, :]
    row = (np.arange(h, dtype=np.float32) * np.float32(size_mult))[None, None, :, None]
    x = (norm * col).reshape(B, C, -1).sum(2)
    y = (norm * row).reshape(B, C, -1).sum(2)
    return np.stack([x, y], 2)
