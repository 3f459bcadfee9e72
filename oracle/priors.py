"""CPU restatement (numpy) of the structure-prior heat-map rendering -- TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this; the
product path (sgtapose_b200/priors.py -> csrc/priors.cu) never does.

Follows the reference's host code that turns re-projected keypoints into the four prior maps
the network consumes:
  sgtapose/utilities.py:943-972    affine_transform_and_clip
  sgtapose/utilities.py:800-824    draw_umich_gaussian (radius 4, sigma 2, integer centre, max-blend)
  sgtapose/utilities.py:846-853    gaussian2D
  sgtapose/utilities.py:1045-1057  get_prev_hm_wo_noise        -> pre_hm / repro_hm   [H, W]
  sgtapose/utilities.py:1085-1098  get_prev_hm_wo_noise_cls    -> *_hm_cls            [7, H/4, W/4]
  sgtapose/utilities.py:889-925    get_affine_transform (cv2.getAffineTransform of 3 point pairs)
Pinned by tests/golden/priors.npz, generated from the unmodified reference functions by
oracle/make_golden_priors.py.
"""
import numpy as np

RADIUS = 4
SIGMA = 2.0


def gaussian_table(radius=RADIUS, sigma=SIGMA):
    """utilities.py:846-853 with res = [0, 0]: float64 [2r+1, 2r+1]."""
    r = np.arange(-radius, radius + 1, dtype=np.float64)
    y, x = r[:, None], r[None, :]
    h = np.exp(-(x * x + y * y) / (2 * sigma * sigma))
    h[h < np.finfo(h.dtype).eps * h.max()] = 0
    return h


def get_affine_transform(center, scale, output_size):
    """utilities.py:889-925 for rot = 0, shift = 0, inv = 0.  The three point pairs are
    (c, c + (0, -s/2), third point), float32, handed to cv2.getAffineTransform like the reference."""
    if not isinstance(scale, (np.ndarray, list)):
        scale = np.array([scale, scale], dtype=np.float32)
    src_w = scale[0]
    dst_w, dst_h = output_size
    src = np.zeros((3, 2), np.float32)
    dst = np.zeros((3, 2), np.float32)
    src_dir = np.array([0, src_w * -0.5], np.float32)        # get_dir([0, -s/2], 0)
    dst_dir = np.array([0, dst_w * -0.5], np.float32)
    src[0] = center
    src[1] = np.asarray(center, np.float32) + src_dir
    dst[0] = [dst_w * 0.5, dst_h * 0.5]
    dst[1] = np.array([dst_w * 0.5, dst_h * 0.5], np.float32) + dst_dir
    for a in (src, dst):
        d = a[0] - a[1]
        a[2] = a[1] + np.array([-d[1], d[0]], np.float32)
    import cv2                      # the reference calls cv2.getAffineTransform (utilities.py:919-922): same solver
    return cv2.getAffineTransform(np.float32(src), np.float32(dst))


def affine_transform_and_clip(pts, t, width, height, raw_width, raw_height):
    """utilities.py:943-972: float64 affine, clip to the map, (0,0) for points outside the raw image."""
    pts = np.asarray(pts, np.float64)
    n = pts.shape[0]
    hom = np.concatenate((pts, np.ones((n, 1))), axis=-1)
    new = np.dot(t, hom.T).T
    new[:, 0] = np.clip(new[:, 0], 0, width - 1)
    new[:, 1] = np.clip(new[:, 1], 0, height - 1)
    inside = (pts[:, 0] >= 0.0) & (pts[:, 0] < raw_width) & (pts[:, 1] >= 0.0) & (pts[:, 1] < raw_height)
    new[~inside] = 0
    return new


def draw_umich_gaussian(hm, center, radius=RADIUS):
    """utilities.py:800-824 (k = 1): nothing is drawn unless the whole patch plus one pixel fits."""
    h, w = hm.shape
    x, y = int(center[0]), int(center[1])
    if x - radius >= 0 and x + radius + 1 < w and y - radius >= 0 and y + radius + 1 < h:
        g = gaussian_table(radius)
        patch = hm[y - radius:y + radius + 1, x - radius:x + radius + 1]
        np.maximum(patch, g, out=patch)
    return hm


def render_hm(centres, h, w):
    """get_prev_hm_wo_noise after the affine: all keypoints max-blended into ONE [h, w] map."""
    hm = np.zeros((h, w), np.float32)
    for c in centres:
        draw_umich_gaussian(hm, c)
    return hm


def render_hm_cls(centres, h, w):
    """get_prev_hm_wo_noise_cls after the affine: one [h, w] map per keypoint."""
    hm = np.zeros((len(centres), h, w), np.float32)
    for i, c in enumerate(centres):
        draw_umich_gaussian(hm[i], c)
    return hm


def get_prev_hm_wo_noise(kp_raw, trans, w, h, raw_w, raw_h):
    if kp_raw is None:
        return np.zeros((h, w), np.float32)
    return render_hm(affine_transform_and_clip(kp_raw, trans, w, h, raw_w, raw_h), h, w)


def get_prev_hm_wo_noise_cls(kp_raw, n_kp, trans, w, h, raw_w, raw_h):
    if kp_raw is None:
        return np.zeros((n_kp, h, w), np.float32)
    return render_hm_cls(affine_transform_and_clip(kp_raw, trans, w, h, raw_w, raw_h), h, w)
