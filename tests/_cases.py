"""Shared, deterministic test-case definitions (inputs are regenerated from seeds on every
box; only reference OUTPUTS are stored under tests/golden/)."""
import numpy as np
import torch

from sgtapose_b200 import synth

GOLDEN_SEED = 317

# (name, B, Cin, Cout, H, W) -- DeformConv = DCN(3x3,s1,p1,d1,dg1) + BN + ReLU as at dla.py:538-550
DEFORMCONV_CASES = [
    ("c64_64_24", 2, 64, 64, 24, 24),
    ("c128_64_16", 1, 128, 64, 16, 16),
    ("c64_128_12x20", 1, 64, 128, 12, 20),
    ("c256_256_6", 1, 256, 256, 6, 6),
]


def gen(seed, *shape, scale=1.0):
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


def deformconv_params(name, Cin, Cout):
    """State-dict of one DeformConv block (keys relative to the block)."""
    s = abs(hash_name(name))
    bound = (1.0 / (Cin * 9)) ** 0.5
    return {
        "conv.weight": gen(s + 1, Cout, Cin, 3, 3) * bound * 1.5,
        "conv.bias": gen(s + 2, Cout) * 0.1,
        "conv.conv_offset_mask.weight": gen(s + 3, 27, Cin, 3, 3) * 0.02,
        "conv.conv_offset_mask.bias": gen(s + 4, 27) * 0.8,
        "actf.0.weight": gen(s + 5, Cout).abs() * 0.5 + 0.5,
        "actf.0.bias": gen(s + 6, Cout) * 0.1,
        "actf.0.running_mean": gen(s + 7, Cout) * 0.1,
        "actf.0.running_var": gen(s + 8, Cout).abs() * 0.5 + 0.5,
        "actf.0.num_batches_tracked": torch.zeros((), dtype=torch.long),
    }


def hash_name(name):
    import zlib
    return zlib.crc32(name.encode()) & 0x7FFFFFF


def deformconv_input(name, B, Cin, H, W):
    return gen(hash_name(name) + 100, B, Cin, H, W)


def decode_heatmaps():
    """[N,7,96,96] heatmaps covering the edge cases of SURVEY.md 8c(4)."""
    hm, _ = synth.synthetic_heatmaps(6, 7, 96, 96, seed=GOLDEN_SEED, noise=0.005, missing_every=5)
    hm = hm.numpy().copy()
    yy, xx = np.mgrid[0:96, 0:96].astype(np.float32)
    blob = lambda cx, cy, a=1.0: (a * np.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / 8.0)).astype(np.float32)
    e = np.zeros((4, 7, 96, 96), np.float32)
    e[0, 1] = blob(30, 40) + blob(70, 20)                  # two equal peaks -> ambiguous
    e[0, 2] = blob(30, 40) + blob(70, 60, 0.5)             # gap 0.5 >= 0.25, sorted by y: lower-y one loses?
    e[0, 3] = blob(30, 60) + blob(70, 20, 0.5)
    e[0, 4] = blob(30, 40) + blob(60, 60, 0.8)             # gap 0.2 < 0.25 -> missing
    e[0, 5] = blob(0, 0)                                   # peak in the corner
    e[0, 6] = blob(95, 50)                                 # peak on the right border
    e[1, 0] = blob(50.5, 50.5)                             # centroid exactly between pixels
    e[1, 1] = blob(10, 94.7)
    e[1, 2][40:43, 40:43] = 1.0                            # plateau
    e[1, 3][:] = 0.02                                      # flat above threshold everywhere
    e[1, 4] = blob(20, 20, 0.05)                           # weak: blurred max below 0.01? (score 0.05)
    e[1, 5] = blob(48, 48) + blob(52, 48)                  # two close blobs merging after blur
    e[1, 6] = blob(5, 5) + blob(90, 90, 0.7) + blob(5, 90, 0.3)
    rng = np.random.default_rng(7)
    e[2] = rng.random((7, 96, 96), dtype=np.float32)       # pure noise: many candidates
    e[3] = rng.random((7, 96, 96), dtype=np.float32) * 0.02
    return np.concatenate([hm, e], 0)


def decode_reg_tracking(n):
    reg = gen(11, n, 2, 96, 96).numpy() * 0.3
    trk = gen(12, n, 2, 96, 96).numpy() * 2.0
    return reg.astype(np.float32), trk.astype(np.float32)


def prior_maps_for_index_cases():
    """[B,7,96,96] prior class maps for the token-index cases incl. the level-3 quirk (47,47)."""
    m = np.zeros((3, 7, 96, 96), np.float32)
    pts = [(47, 47), (0, 0), (95, 95), (3, 90), (50, 1), (94, 2), (33, 77)]
    for c, (x, y) in enumerate(pts):
        m[0, c, y, x] = 1.0
    rng = np.random.default_rng(3)
    m[1] = rng.random((7, 96, 96), dtype=np.float32)
    # sample 2 stays all-zero: frame-0 behaviour (every top-1 index is 0)
    return m


# ----------------------------------------------------------------------------- host PnP / pose / metrics cases
MISSING = -999.999 * 4
CAMERA_K = np.array([[502.30, 0.0, 319.75], [0.0, 502.30, 179.75], [0.0, 0.0, 1.0]])    # sgta_detector.py:83
RAW_W, RAW_H = 640, 360


def panda_scene(rng, n):
    """[n,7,3] Panda-like keypoint positions in front of the camera (metres, camera frame)."""
    return rng.uniform([-0.35, -0.2, 1.2], [0.35, 0.2, 1.8], size=(n, 7, 3))


def project(x3d, K=CAMERA_K):
    p = np.einsum("ij,...j->...i", K, x3d)
    return p[..., :2] / p[..., 2:]


def pnp_cases():
    """(prev_pos [7,3], detected kps [7,2] with MISSING rows, next_pos [7,3]) of geometric_vision.is_pnp."""
    rng = np.random.default_rng(41)
    cases = []
    for i in range(8):
        prev = panda_scene(rng, 1)[0]
        nxt = prev + rng.normal(0, 0.004, size=prev.shape)
        kps = project(prev) + rng.normal(0, [0.0, 0.3, 1.0, 3.0][i % 4], size=(7, 2))
        if i in (2, 5):
            kps[[1, 4]] = MISSING
        if i == 6:
            kps[[0, 1, 2]] = MISSING                  # four detections left: the EPnP minimum
        if i == 7:
            kps[[0, 1, 2, 3, 5]] = MISSING            # two detections: cv2 raises, the reference swallows it
        cases.append((prev, kps, nxt))
    return cases


def pose_heatmap_cases(n=12, q=96):
    """Synthetic head outputs whose blobs sit at the projections of a known scene:
    hm [n,7,q,q] (post-sigmoid range), reg, tracking [n,2,q,q], x3d [n,7,3] = the keypoints in the ROBOT frame
    (a random rigid transform away from the camera frame, so the recovered pose is a real rotation + translation)."""
    import cv2
    from sgtapose_b200 import priors as PR
    rng = np.random.default_rng(43)
    x3d_cam = panda_scene(rng, n)
    x3d = np.empty_like(x3d_cam)
    for i in range(n):
        R0 = cv2.Rodrigues(rng.normal(0, 0.9, size=3))[0]
        t0 = rng.uniform([-0.2, -0.2, 0.8], [0.2, 0.2, 1.4])
        x3d[i] = (x3d_cam[i] - t0) @ R0                       # rows: R0^T (x_cam - t0)
    c = np.array([RAW_W / 2.0, RAW_H / 2.0], dtype=np.float32)
    trans_out = PR.get_affine_transform(c, max(RAW_H, RAW_W) * 1.0, 0, [q, q])
    raw = project(x3d_cam)
    ctr = np.einsum("ij,nkj->nki", trans_out, np.concatenate([raw, np.ones((n, 7, 1))], 2))
    yy, xx = np.mgrid[0:q, 0:q].astype(np.float32)
    hm = rng.random((n, 7, q, q), dtype=np.float32) * np.float32(0.004)
    for i in range(n):
        for k in range(7):
            if i % 5 == 4 and k == i % 7:
                continue                                # a keypoint the network missed
            cx, cy = ctr[i, k]
            hm[i, k] += (0.9 * np.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / 8.0)).astype(np.float32)
    reg = (rng.random((n, 2, q, q), dtype=np.float32) * np.float32(0.8)).astype(np.float32)
    trk = rng.normal(0, 1.5, size=(n, 2, q, q)).astype(np.float32)
    return np.clip(hm, 0, 1), reg, trk, x3d


def metrics_case():
    rng = np.random.default_rng(47)
    n = 400
    gt = rng.uniform([-60, -40], [700, 400], size=(n, 2))
    det = gt + rng.normal(0, 4.0, size=(n, 2))
    det[rng.random(n) < 0.15] = MISSING
    add = np.abs(rng.normal(0.02, 0.02, size=150))
    add[rng.random(150) < 0.1] = -999.99
    inframe = rng.integers(2, 8, size=150)
    return det, gt, add, inframe
