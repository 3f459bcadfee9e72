"""Lock-step clip runner (sgtapose_b200/detector.py) vs the per-clip restatement of the reference's
host steps (oracle/detector.py) and the oracle's decode -- GPU tests."""
import numpy as np
import pytest
import torch

from oracle import decode as odec
from oracle import detector as odet
from tests import _cases as C

pytestmark = pytest.mark.gpu
DEV = "cuda"
S, B = 128, 3


def _scene(rng, B, frame):
    """7 Panda-like keypoints in front of the camera, drifting a little per frame: [B,7,3]."""
    base = rng.uniform([-0.35, -0.2, 1.2], [0.35, 0.2, 1.8], size=(B, 7, 3))
    return base + 0.004 * frame


@pytest.fixture(scope="module")
def det():
    from sgtapose_b200 import config, detector, engine, networks, synth
    m = networks.create_model(config.ARCH, dict(config.HEADS), dict(config.HEAD_CONV), config.default_opt())
    sd = synth.synthetic_state_dict(m.state_dict(), seed=C.GOLDEN_SEED)
    eng = engine.InferenceEngine(sd, config.default_opt(), batch=B, size=S, mode="fp32", device=DEV, fuse_sigmoid=True)
    return detector.LockstepDetector(eng, workers=4)


def test_post_process_matches_reference_loops(det):
    rng = np.random.default_rng(1)
    scores = rng.uniform(-1, 1, size=(B, 7)).astype(np.float32)
    scores[0, 2] = -1.0
    scores[1] = 0.0005                                   # below out_thresh = 0.001
    cts = rng.uniform(0, S // 4, size=(B, 7, 2)).astype(np.float32)
    from sgtapose_b200 import detector
    got = detector.post_process_device(torch.from_numpy(scores).to(DEV), torch.from_numpy(cts).to(DEV), det.trans_inv,
                                       det.out_thresh).cpu().numpy()
    assert np.array_equal(got, detector.post_process_batch(scores, cts, det.trans_inv, det.out_thresh))
    for b in range(B):
        want = odet.final_kps(odet.post_process_one(scores[b], cts[b], det.trans_inv, det.out_thresh), 7)
        assert np.array_equal(got[b], want), b


def test_step_sequence_priors_and_decode(det):
    """Three frames of three clips.  Every frame: (1) the prior maps the runner rendered into the engine's
    input buffers == the oracle's host rendering from the SAME detections, through the ORACLE's own PnP front-end
    (oracle/detector.py::is_pnp, pinned to the reference in tests/golden/pnp.npz), bit-exact; (2) the
    detections it reports == reference post-processing of the oracle's decode of the engine's heads."""
    from sgtapose_b200 import detector, synth
    rng = np.random.default_rng(2)
    det.reset()
    q = S // 4
    x3d = [_scene(np.random.default_rng(7), B, f) for f in range(3)]
    for f in range(3):
        imgs = synth.synthetic_inputs(B, S, seed=100 + f, frame=1)[0]
        if f == 1:
            # plant detections so that the PnP branch runs: exact projections for clip 0, two missing
            # keypoints for clip 1, nothing detected for clip 2
            proj = np.einsum("ij,bkj->bki", det.K, x3d[0])
            proj = proj[:, :, :2] / proj[:, :, 2:]
            kps = proj.copy()
            kps[1, [2, 5]] = odet.MISSING
            kps[2] = odet.MISSING
            det.detected_kps = kps
        before = det.detected_kps.copy()
        out = det.step(imgs.pin_memory(), x3d[f - 1] if f else None, x3d[f] if f else None)
        inp = det.eng.inp
        for b in range(B):
            if f == 0:
                want = (np.zeros((S, S), np.float32),) * 2 + (np.zeros((7, q, q), np.float32),) * 2
            else:
                want = odet.further_inputs(before[b], x3d[f - 1][b], x3d[f][b], det.K, det.trans_input,
                                           det.trans_output, S, q, det.raw_w, det.raw_h)
            got = (inp["pre_hm"][b, 0], inp["repro_hm"][b, 0], inp["pre_hm_cls"][b], inp["repro_hm_cls"][b])
            for name, g, w in zip(("pre_hm", "repro_hm", "pre_hm_cls", "repro_hm_cls"), got, want):
                assert np.array_equal(g.cpu().numpy(), w), (f, b, name)
            if f == 1 and b == 0:
                assert inp["pre_hm"][b].max().item() == 1.0 and inp["repro_hm_cls"][b].sum().item() > 7
        heads = det.eng.out
        ref = odec.dream_generic_decode(heads["hm"].cpu().numpy(), heads["reg"].cpu().numpy(),
                                        heads["tracking"].cpu().numpy())
        for b in range(B):
            want = odet.final_kps(odet.post_process_one(np.asarray(ref["scores"][b], np.float32),
                                                        np.asarray(ref["cts_wreg"][b], np.float32).reshape(7, 2),
                                                        det.trans_inv, det.out_thresh), 7)
            assert np.array_equal(out["kps_raw"][b], want), (f, b)
        # previous image handed over on the device
        if f:
            assert torch.equal(inp["pre_img"], prev_x)
        prev_x = inp["x"].clone()
    assert det.frame == 3 and det.timing["steps"] == 3


def test_post_process_device_vs_oracle():
    """sgta_post_process (post_process + merge_outputs + _get_final_kps on the device) vs the per-item restatement of
    post_process.py:93-117 / sgta_detector.py:608-651, :955-961 -- bit-exact, incl. scores at and around the threshold."""
    from sgtapose_b200 import detector
    from sgtapose_b200 import priors as PR
    rng = np.random.default_rng(3)
    Bn, K = 37, 7
    scores = rng.random((Bn, K)).astype(np.float32)
    scores[0, :3] = [0.001, np.float32(0.001) + np.float32(1e-9), 0.00099]
    scores[1] = -1.0                                            # the decode's "missing" score
    cts = (rng.random((Bn, K, 2)) * 96).astype(np.float32)
    c, s = np.array([320.0, 180.0], np.float32), 640.0
    trans_inv = PR.get_affine_transform(c, s, 0, (96, 96), inv=1).astype(np.float32)
    thresh = 0.001
    got = detector.post_process_device(torch.from_numpy(scores).to(DEV), torch.from_numpy(cts).to(DEV), trans_inv, thresh).cpu().numpy()
    for b in range(Bn):
        want = odet.final_kps(odet.post_process_one(scores[b], cts[b], trans_inv, thresh), K)
        assert np.array_equal(got[b], want), b
    assert np.array_equal(got, detector.post_process_batch(scores, cts, trans_inv, thresh))   # the host form agrees


def test_step_accepts_raw_uint8_frames(det):
    """Raw 640x360 uint8 frames pre-processed on the device == the oracle's pre_process of the same frames
    fed as float32 network inputs: identical input buffer and identical detections."""
    from oracle import preprocess as opre
    from oracle.make_golden_preprocess import case_image
    frames = np.stack([case_image((det.raw_h, det.raw_w), 40 + b) for b in range(B)])
    want_x = np.concatenate([opre.pre_process(frames[b], S, S)[0] for b in range(B)])
    det.reset()
    out_raw = det.step(frames)
    assert np.array_equal(det.eng.inp["x"].cpu().numpy(), want_x)
    assert torch.equal(det.eng.inp["pre_img"], det.eng.inp["x"])
    det.reset()
    out_f32 = det.step(torch.from_numpy(want_x).pin_memory())
    assert np.array_equal(out_raw["kps_raw"], out_f32["kps_raw"]) and np.array_equal(out_raw["scores"], out_f32["scores"])
    with pytest.raises(ValueError):
        det.step(frames[:, :100])


def test_clip_groups_skewed_schedule_equals_group_by_group(det):
    """Two groups (2 + 1 clips, one engine each) run skewed == the same two detectors stepped one after the
    other: the schedule only re-orders host work, every frame's detections are identical, with the PnP branch
    active from frame 1 on.  (Engines of different batch size are NOT compared bit-wise: reduction splits in
    the token kernels depend on the token count, and the synthetic weights amplify 1-ulp differences.)"""
    from sgtapose_b200 import config, detector, engine, networks, synth
    m = networks.create_model(config.ARCH, dict(config.HEADS), dict(config.HEAD_CONV), config.default_opt())
    sd = synth.synthetic_state_dict(m.state_dict(), seed=C.GOLDEN_SEED)
    engs = [engine.InferenceEngine(sd, config.default_opt(), batch=b, size=S, mode="fp32", device=DEV, fuse_sigmoid=True)
            for b in (2, 1)]
    groups = detector.ClipGroups([detector.LockstepDetector(e, workers=2) for e in engs])
    assert groups.B == B
    n_frames = 4
    x3d = [_scene(np.random.default_rng(7), B, f) for f in range(n_frames)]
    imgs = [synth.synthetic_inputs(B, S, seed=200 + f, frame=1)[0].pin_memory() for f in range(n_frames)]
    proj = np.einsum("ij,bkj->bki", det.K, x3d[0])
    planted = proj[:, :, :2] / proj[:, :, 2:]
    planted[1, [2, 5]] = odet.MISSING
    off = groups.offsets

    # reference run: each group on its own, plain step()
    want = [dict() for _ in range(n_frames)]
    for g, d in enumerate(groups.dets):
        d.reset()
        for f in range(n_frames):
            if f == 1:
                d.detected_kps = planted[off[g]:off[g + 1]].copy()
            sl = slice(off[g], off[g + 1])
            want[f][g] = d.step(imgs[f][sl], x3d[f - 1][sl] if f else None, x3d[f][sl] if f else None)

    def plant(g, f, d):
        if f == 1:
            d.detected_kps = planted[off[g]:off[g + 1]].copy()

    groups.reset()
    got = groups.run(n_frames, lambda f: imgs[f], lambda f: x3d[f], before_begin=plant)
    for f in range(n_frames):
        for k in ("kps_raw", "scores"):
            assert np.array_equal(got[f][k], np.concatenate([want[f][0][k], want[f][1][k]])), (f, k)
    assert all(d.frame == n_frames for d in groups.dets)
    with pytest.raises(RuntimeError):
        groups.dets[0].finish()                            # finish() without begin()
