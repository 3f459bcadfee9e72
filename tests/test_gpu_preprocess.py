"""Device pre-processing (sgtapose_b200/preprocess.py, csrc/preprocess.cu) vs the reference's
pre_process output (golden), the oracle restatement and cv2 itself -- GPU tests, bit-exact."""
import types

import numpy as np
import pytest
import torch

from oracle import preprocess as opre
from oracle.make_golden_preprocess import CASES, case_image

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_pre_process_matches_reference_golden(golden):
    from sgtapose_b200 import preprocess as pre
    g = golden("preprocess.npz")
    for i, (raw, inp, seed) in enumerate(CASES):
        opt = types.SimpleNamespace(fix_res=True, fix_short=-1, input_h=inp[0], input_w=inp[1], down_ratio=4)
        images, meta = pre.pre_process(case_image(raw, seed), opt, device=DEV)
        assert images.shape == (1, 3, inp[0], inp[1]) and images.dtype == torch.float32
        assert np.array_equal(images.cpu().numpy(), g["images_%d" % i]), i      # float32 bit-exact
        assert np.array_equal(meta["trans_input"], g["trans_input_%d" % i])


def test_warp_u8_matches_cv2_and_oracle():
    import cv2
    from sgtapose_b200 import preprocess as pre
    rng = np.random.default_rng(7)
    B = 70                                                   # > 64: two launches, per-frame matrices
    frames = rng.integers(0, 256, (B, 45, 80, 3), dtype=np.uint8)
    mats = np.zeros((B, 2, 3))
    for b in range(B):
        a, s = rng.uniform(-0.6, 0.6), rng.uniform(0.4, 2.2)
        mats[b] = [[s * np.cos(a), -s * np.sin(a) + rng.uniform(-0.1, 0.1), rng.uniform(-30, 30)],
                   [s * np.sin(a), s * np.cos(a), rng.uniform(-30, 30)]]
    H, W = 52, 77
    out, u8 = pre.warp_normalize(torch.from_numpy(frames).to(DEV), mats, (H, W), return_u8=True)
    u8 = u8.cpu().numpy()
    for b in range(B):
        ref = cv2.warpAffine(frames[b], mats[b], (W, H), flags=cv2.INTER_LINEAR)
        assert np.array_equal(u8[b], ref), b
        if b < 6:
            assert np.array_equal(u8[b], opre.warp_affine_u8(frames[b], mats[b], (W, H)))
    mean = np.array([0.5, 0.5, 0.5], np.float32).reshape(1, 1, 3)
    want = np.stack([opre.normalize(u8[b], mean, mean).transpose(2, 0, 1) for b in range(B)])
    assert np.array_equal(out.cpu().numpy(), want)
    # one shared matrix for the whole batch == per-frame copies of it
    a = pre.warp_normalize(torch.from_numpy(frames).to(DEV), mats[3], (H, W))
    bb = pre.warp_normalize(torch.from_numpy(frames).to(DEV), np.repeat(mats[3:4], B, 0), (H, W))
    assert torch.equal(a, bb)


def test_pre_process_bench_shape_batch():
    """BASELINE frame size at the bench batch: 32 raw 640x360 frames -> [32,3,384,384], equal to the oracle per frame."""
    from sgtapose_b200 import preprocess as pre
    opt = types.SimpleNamespace(fix_res=True, fix_short=-1, input_h=384, input_w=384, down_ratio=4)
    frames = np.stack([case_image((360, 640), 100 + b) for b in range(32)])
    images, meta = pre.pre_process(frames, opt, device=DEV)
    got = images.cpu().numpy()
    for b in (0, 13, 31):
        want, _, _ = opre.pre_process(frames[b], 384, 384)
        assert np.array_equal(got[b:b + 1], want)
