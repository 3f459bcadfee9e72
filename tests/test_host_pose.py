"""Host PnP front-end, pose-level parity (north_star: poses within 1 mm / 0.1 deg of the reference) and the offline
metrics, against goldens produced by the UNMODIFIED reference functions (oracle/make_golden_r2.py; pyrr stubbed as
oracle/ref_geom.py states).  CPU tests use the oracle's decode; the GPU test feeds the CUDA decode into the same
host code.  Nothing here passes a product function into an oracle."""
import numpy as np
import pytest
import torch

from tests import _cases as C

POS_TOL_M = 1e-3           # 1 mm
ROT_TOL_DEG = 0.1


def _rot_angle_deg(q1, q2):
    d = abs(float(np.dot(q1 / np.linalg.norm(q1), q2 / np.linalg.norm(q2))))
    return np.degrees(2.0 * np.arccos(min(1.0, d)))


def test_is_pnp_and_solve_pnp_vs_reference_golden(golden):
    from sgtapose_b200 import detector
    g = golden("pnp.npz")
    for i, (prev, kps, nxt) in enumerate(C.pnp_cases()):
        good = np.unique(np.where(kps > C.MISSING)[0])
        a, b = detector.is_pnp(prev[good], kps[good], nxt, kps, C.CAMERA_K)
        assert np.array_equal(a, g["prev_%d" % i]), i
        assert np.allclose(b, g["next_%d" % i], rtol=0, atol=1e-7), (i, np.abs(b - g["next_%d" % i]).max())
        ok, t, q = detector.solve_pnp_quat(prev[good], kps[good], C.CAMERA_K)
        assert ok == bool(g["ok_%d" % i]), i
        if ok:
            assert np.allclose(t, g["t_%d" % i], rtol=0, atol=1e-9) and np.allclose(q, g["q_%d" % i], rtol=0, atol=1e-9)
            assert np.allclose(detector.rotation_from_quaternion(q), g["R_%d" % i], rtol=0, atol=1e-9)
    # the un-solvable case (two detections) falls back to the detections themselves (geometric_vision.py:308-310)
    prev, kps, nxt = C.pnp_cases()[7]
    good = np.unique(np.where(kps > C.MISSING)[0])
    a, b = detector.is_pnp(prev[good], kps[good], nxt, kps, C.CAMERA_K)
    assert a is kps and b is kps


def test_rotation_restatement_vs_cv2_rodrigues():
    """The pyrr restatement (quaternion_from_rvec -> rotation_from_quaternion) is the rotation cv2 itself assigns to
    the same rvec: an independent check of the one piece of the PnP path that is restated rather than imported."""
    import cv2
    from sgtapose_b200 import detector
    rng = np.random.default_rng(3)
    for _ in range(50):
        rvec = rng.normal(0, 1.2, size=3)
        assert np.allclose(detector._rotation_from_rvec(rvec), cv2.Rodrigues(rvec)[0], rtol=0, atol=1e-12)


def _poses_from_decode(scores, cts_wreg, x3d):
    from sgtapose_b200 import detector
    from sgtapose_b200 import priors as PR
    q = 96
    c = np.array([C.RAW_W / 2.0, C.RAW_H / 2.0], dtype=np.float32)
    trans_inv = PR.get_affine_transform(c, max(C.RAW_H, C.RAW_W) * 1.0, 0, (q, q), inv=1).astype(np.float32)
    kps = detector.post_process_batch(np.asarray(scores, np.float32), np.asarray(cts_wreg, np.float32).reshape(-1, 7, 2),
                                      trans_inv, 0.001)
    poses, oks = [], []
    for n in range(kps.shape[0]):
        good = np.unique(np.where(kps[n] > -999.0)[0])
        ok, t, quat = detector.solve_pnp_quat(x3d[n][good], kps[n][good], C.CAMERA_K)
        oks.append(ok)
        poses.append(np.concatenate([t, quat]) if ok else np.full(7, -999.99))
    return kps, np.stack(poses), np.array(oks)


def _check_poses(kps, poses, oks, g):
    assert np.array_equal(oks, g["ok"])
    assert np.array_equal(kps, g["kps_raw"])                       # same float32 affine, same selection: exact
    for n in range(len(oks)):
        dt = np.linalg.norm(poses[n, :3] - g["pose_xyz_xyzw"][n, :3])
        dr = _rot_angle_deg(poses[n, 3:], g["pose_xyz_xyzw"][n, 3:])
        assert dt <= POS_TOL_M and dr <= ROT_TOL_DEG, (n, dt, dr)


def test_pose_parity_oracle_decode_vs_reference_pipeline(golden):
    """heat maps -> oracle decode -> product post-process + PnP  ==  reference decode -> post_process ->
    _get_final_kps -> solve_pnp (pose.npz), within 1 mm / 0.1 deg (in fact ~1e-12)."""
    from oracle import decode as odec
    hm, reg, trk, x3d = C.pose_heatmap_cases()
    d = odec.dream_generic_decode(hm, reg, trk)
    _check_poses(*_poses_from_decode(d["scores"], d["cts_wreg"], x3d), golden("pose.npz"))


@pytest.mark.gpu
def test_pose_parity_gpu_decode_vs_reference_pipeline(golden):
    """The same with the CUDA live decode (sgta_decode_peaks through sgtapose_b200.decode) producing the keypoints."""
    from sgtapose_b200 import config, decode
    hm, reg, trk, x3d = C.pose_heatmap_cases()
    out = {"hm": torch.from_numpy(hm).cuda(), "reg": torch.from_numpy(reg).cuda(), "tracking": torch.from_numpy(trk).cuda()}
    d = decode.dream_generic_decode(out, K=7, opt=config.default_opt())
    _check_poses(*_poses_from_decode(d["scores"].cpu().numpy(), d["cts_wreg"].cpu().numpy(), x3d), golden("pose.npz"))


def test_metrics_vs_reference_golden(golden):
    from sgtapose_b200 import metrics
    g = golden("metrics.npz")
    det, gt, add, inframe = C.metrics_case()
    for syn in (False, True):
        r = metrics.keypoint_metrics(det, gt, None, (C.RAW_W, C.RAW_H), 12.0, syn)
        for k, v in r.items():
            want = float(g["kp_%d_%s" % (syn, k)])
            assert v == pytest.approx(want, rel=1e-12, abs=1e-12), (syn, k, v, want)
    r = metrics.pnp_metrics(add, inframe)
    for k, v in r.items():
        assert float(v) == pytest.approx(float(g["pnp_%s" % k]), rel=1e-12, abs=1e-12), k
    rng = np.random.default_rng(5)
    x3d = C.panda_scene(rng, 4)
    for i in range(4):
        rng.normal(0, 1.0, size=(7, 2))                        # keep the generator in step with the fixture script
        a = metrics.add_from_pose(g["add_t_%d" % i], g["add_q_%d" % i], x3d[i])
        assert a == pytest.approx(float(g["add_values"][i]), rel=1e-12)


@pytest.mark.reference
def test_host_pnp_vs_live_reference():
    """With /root/reference present: the reference's own is_pnp / solve_pnp on fresh random cases."""
    from oracle import ref_geom
    from sgtapose_b200 import detector
    gv = ref_geom.load_geometric_vision()
    rng = np.random.default_rng(99)
    for i in range(20):
        prev = C.panda_scene(rng, 1)[0]
        nxt = prev + rng.normal(0, 0.004, size=prev.shape)
        kps = C.project(prev) + rng.normal(0, 1.0, size=(7, 2))
        kps[rng.random(7) < 0.2] = C.MISSING
        good = np.unique(np.where(kps > C.MISSING)[0])
        if len(good) < 4:
            continue
        wa, wb = gv.is_pnp(prev[good], kps[good], nxt, kps, C.CAMERA_K)
        ga, gb = detector.is_pnp(prev[good], kps[good], nxt, kps, C.CAMERA_K)
        assert np.array_equal(wa, ga) and np.allclose(wb, gb, rtol=0, atol=1e-9)
        ok, t, q = gv.solve_pnp(prev[good], kps[good], C.CAMERA_K)
        ok2, t2, q2 = detector.solve_pnp_quat(prev[good], kps[good], C.CAMERA_K)
        assert bool(ok) == ok2 and np.allclose(t, t2, atol=1e-12) and np.allclose(np.asarray(q), q2, atol=1e-12)
