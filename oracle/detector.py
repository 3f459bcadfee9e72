"""CPU restatement of the reference's per-clip host steps around the network -- TEST INFRASTRUCTURE.

Only tests/ may import this.  Per-item loops, written like the reference:
  dream_generic_post_process   sgtapose/lib/utils/post_process.py:93-117 (ct_wreg, score, class only)
  merge_outputs                sgtapose/lib/sgta_detector.py:955-961
  _get_final_kps               sgtapose/lib/sgta_detector.py:608-651 (is_ct branch)
  _get_further_dt_pnp_inputs_real   :501-547 (x3d given instead of read from JSON), rendering via oracle/priors.py
  is_pnp / solve_pnp           sgtapose/geometric_vision.py:283-310, :43-116 -- the rotation comes from cv2.Rodrigues
                               here (the reference goes rvec -> pyrr quaternion -> matrix33; pyrr is absent), pinned to
                               the reference functions' own outputs in tests/golden/pnp.npz (tests/test_oracle.py)
"""
import numpy as np

from . import priors as OP

MISSING = -999.999 * 4


def post_process_one(scores, cts_wreg, trans_inv_f32, out_thresh):
    """scores [K], cts_wreg [K,2] of ONE clip -> list of {'score','class','ct_wreg'} (post_process + merge)."""
    preds = []
    for j in range(len(scores)):
        if scores[j] < out_thresh:
            continue
        tc = np.ones((1, 3), np.float32)
        tc[:, :2] = cts_wreg[j].reshape(1, 2)
        ct = np.dot(trans_inv_f32, tc.transpose()).transpose()[:, :2].reshape(2)
        preds.append({"score": scores[j], "class": j + 1, "ct_wreg": ct})
    return [p for p in preds if p["score"] > out_thresh]


def final_kps(dets, num_classes):
    out = np.full((num_classes, 2), MISSING)
    cls = {i: {"x": [], "y": []} for i in range(1, num_classes + 1)}
    for det in dets:
        ct = det["ct_wreg"].tolist()
        cls[det["class"]]["x"].append([det["score"], ct[0]])
        cls[det["class"]]["y"].append([det["score"], ct[1]])
    for i in range(1, num_classes + 1):
        xl, yl = cls[i]["x"], cls[i]["y"]
        if not xl:
            continue
        xl.sort(); yl.sort()
        out[i - 1] = [xl[-1][1], yl[-1][1]]
    return out


def solve_pnp(points, projs, camera_K):
    import cv2
    if len(points) == 0:
        return False, None, None
    try:
        obj, img = np.asarray(points, np.float64).reshape(-1, 1, 3), np.asarray(projs, np.float64).reshape(-1, 1, 2)
        ok, rvec, tvec = cv2.solvePnP(obj, img, camera_K, np.array([]), flags=cv2.SOLVEPNP_EPNP)
        ok, rvec, tvec = cv2.solvePnP(obj, img, camera_K, np.array([]), flags=cv2.SOLVEPNP_ITERATIVE,
                                      useExtrinsicGuess=True, rvec=rvec, tvec=tvec)
        return ok, tvec[:, 0], cv2.Rodrigues(rvec)[0]
    except Exception:
        return False, None, None


def is_pnp(prev_pos, prev_projs, next_pos, prev_projs_all, camera_K):
    ok, t, R = solve_pnp(prev_pos, prev_projs, camera_K)
    if not ok:
        return prev_projs_all, prev_projs_all
    est = []
    for x in next_pos:                                   # one point at a time, like point_projection_from_3d
        pc = camera_K @ (R @ x + t)
        est.append([pc[0] / pc[2], pc[1] / pc[2]])
    return prev_projs_all, np.array(est)


def further_inputs(kps_detected, x3d_prev, x3d_next, camera_K, trans_input, trans_output, S, q, raw_w, raw_h):
    """-> (pre_hm [S,S], repro_hm [S,S], pre_hm_cls [K,q,q], repro_hm_cls [K,q,q]) of ONE clip."""
    n_kp = x3d_prev.shape[0]
    good = np.unique(np.where(kps_detected > MISSING)[0])
    if len(good) == 0:
        z, zc = np.zeros((S, S), np.float32), np.zeros((n_kp, q, q), np.float32)
        return z, z.copy(), zc, zc.copy()
    prev, nxt = is_pnp(x3d_prev[good], kps_detected[good], x3d_next, kps_detected, camera_K)
    return (OP.get_prev_hm_wo_noise(prev, trans_input, S, S, raw_w, raw_h),
            OP.get_prev_hm_wo_noise(nxt, trans_input, S, S, raw_w, raw_h),
            OP.get_prev_hm_wo_noise_cls(prev, n_kp, trans_output, q, q, raw_w, raw_h),
            OP.get_prev_hm_wo_noise_cls(nxt, n_kp, trans_output, q, q, raw_w, raw_h))
