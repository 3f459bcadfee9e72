"""Offline pose / keypoint metrics around the host PnP (second half) -- host code.

Same names, arguments and result dictionaries as the reference's
  `keypoint_metrics`   sgtapose/analysis.py:1640-1739   (PCK AUC over in-frame keypoints, 0.01 px steps),
  `pnp_metrics`        sgtapose/analysis.py:1742-1793   (ADD statistics and ADD AUC, 1e-5 m steps),
  `add_from_pose`      sgtapose/geometric_vision.py:186-207 (mean 3-D keypoint distance under the estimated pose),
restated on whole arrays: the reference counts `len(np.where(err < v)[0])` once per threshold (1200 and 6000
passes over the data); a sorted error vector and one `searchsorted` give the same integer counts.
Pinned to the reference functions' own outputs: tests/golden/metrics.npz (oracle/make_golden_r2.py).
"""
import numpy as np

from .detector import rotation_from_quaternion


def keypoint_metrics(keypoints_detected, keypoints_gt, all_json_np=None, image_resolution=(640, 360),
                     auc_pixel_threshold=12.0, syn=False):
    det = np.asarray(keypoints_detected, dtype=np.float64).reshape(-1, 2)
    gt = np.asarray(keypoints_gt, dtype=np.float64).reshape(-1, 2)
    gap = 140 if syn else 0
    out = (gt[:, 0] < 0.0 + gap) | (gt[:, 0] > image_resolution[0] - gap) | (gt[:, 1] < 0.0) | (gt[:, 1] > image_resolution[1])
    missing = (det[:, 0] < -999.0) & (det[:, 1] < -999.0)
    num_gt_outframe = int(out.sum())
    num_gt_inframe = int((~out).sum())
    found_in = ~out & ~missing
    kp_errors = det[found_in] - gt[found_in]
    if len(kp_errors) > 0:
        l2 = np.linalg.norm(kp_errors, axis=1)
        mean, median, std = np.mean(l2), np.median(l2), np.std(l2)
        delta_pixel = 0.01
        pck_values = np.arange(0, auc_pixel_threshold, delta_pixel)
        y_values = np.searchsorted(np.sort(l2), pck_values, side="left")          # count of errors < value
        auc = np.trapz(y_values, dx=delta_pixel) / float(auc_pixel_threshold) / float(num_gt_inframe)
    else:
        mean = median = std = auc = None
    return {
        "num_gt_outframe": num_gt_outframe,
        "num_missing_gt_outframe": int((out & missing).sum()),
        "num_found_gt_outframe": int((out & ~missing).sum()),
        "num_gt_inframe": num_gt_inframe,
        "num_found_gt_inframe": int(found_in.sum()),
        "num_missing_gt_inframe": int((~out & missing).sum()),
        "l2_error_mean_px": mean,
        "l2_error_median_px": median,
        "l2_error_std_px": std,
        "l2_error_auc": auc,
        "l2_error_auc_thresh_px": auc_pixel_threshold,
    }


def pnp_metrics(pnp_add, num_inframe_projs_gt, num_min_inframe_projs_gt_for_pnp=4, add_auc_threshold=0.06,
                pnp_magic_number=-999.0):
    pnp_add = np.array(pnp_add)
    num_inframe_projs_gt = np.array(num_inframe_projs_gt)
    found = pnp_add[pnp_add > pnp_magic_number]
    num_pnp_possible = int((num_inframe_projs_gt >= num_min_inframe_projs_gt_for_pnp).sum())
    delta_threshold = 0.00001
    values = np.arange(0.0, add_auc_threshold, delta_threshold)
    counts = np.searchsorted(np.sort(found), values, side="right") / float(num_pnp_possible)   # count of ADD <= value
    return {
        "num_pnp_found": len(found),
        "num_pnp_not_found": num_pnp_possible - len(found),
        "num_pnp_possible": num_pnp_possible,
        "num_min_inframe_projs_gt_for_pnp": num_min_inframe_projs_gt_for_pnp,
        "pnp_magic_number": pnp_magic_number,
        "add_mean": np.mean(found),
        "add_median": np.median(found),
        "add_std": np.std(found),
        "add_max": np.max(found),
        "add_min": np.min(found),
        "add_auc": np.trapz(counts, dx=delta_threshold) / float(add_auc_threshold),
        "add_auc_thresh": add_auc_threshold,
    }


def add_from_pose(translation, quaternion_xyzw, keypoint_positions_wrt_cam_gt, camera_K=None):
    T = np.eye(4)
    T[:3, :3] = rotation_from_quaternion(quaternion_xyzw)
    T[:3, -1] = translation
    gt = np.asarray(keypoint_positions_wrt_cam_gt, dtype=np.float64)
    hom = np.hstack((gt, np.ones((gt.shape[0], 1))))
    aligned = np.transpose(np.matmul(T, np.transpose(hom)))[:, :3]
    return np.mean(np.linalg.norm(aligned - gt, axis=1))
