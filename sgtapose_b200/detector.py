"""Lock-step clip runner: the per-frame loop of the reference detector for B clips at once.

Mirrors `SGTADetector.run` / `process` / `post_process` / `_get_final_kps`
(sgtapose/lib/sgta_detector.py:117-236, :881-927, :929-942, :608-651) and the PnP front-end
`_get_further_dt_pnp_inputs_real` -> `geometric_vision.is_pnp` -> `solve_pnp`
(sgta_detector.py:501-547, sgtapose/geometric_vision.py:43-116, :283-310) with these differences,
all about WHERE things run, none about what is computed:

  * B independent clips advance one frame per `step` (clips are independent, frames serial:
    SURVEY.md 3.1); the network + decode run once for the whole batch (engine.InferenceEngine);
  * the four prior maps are rendered on the device from 2 x 7 centres per clip (priors.py)
    instead of on the host followed by four uploads; the previous frame's image stays on the
    device (the reference keeps `self.pre_images` there too, sgta_detector.py:204);
  * decoded detections come back in ONE device->host copy per step ([B,7] scores + [B,7,2]
    centres) instead of ~7 per clip;
  * the per-clip PnP (cv2, host, as north_star wants) runs on a thread pool -- cv2 releases the GIL.

Keypoint 3-D positions w.r.t. the camera for the previous / current frame are inputs (the
reference reads them from the frame JSONs, sgta_detector.py:511-517).  `step` takes either
network-input images (float32) or the RAW uint8 frames: those are uploaded as uint8 and pre-processed
on the device (cv2 warpAffine + normalise, :368-399 -> preprocess.py, bit-exact) straight into the
engine's input buffer.
"""
import concurrent.futures as cf
import time

import numpy as np
import torch

from . import preprocess as PP
from . import priors as PR

MISSING = -999.999 * 4          # sgta_detector.py:613, :521
DEFAULT_K = np.array([[502.30, 0.0, 319.75], [0.0, 502.30, 179.75], [0.0, 0.0, 1.0]])   # sgta_detector.py:83


def quaternion_from_rvec(rvec):
    """geometric_vision.py:15-25 convert_rvec_to_quaternion: axis-angle -> xyzw quaternion
    (pyrr `Quaternion.from_axis_rotation`, then `.normalize()`).  pyrr is an un-pinned, absent dependency: restated
    from its published formulas (pyrr 0.10.3 quaternion.create_from_axis_rotation / normalize)."""
    theta = np.sqrt(rvec[0] * rvec[0] + rvec[1] * rvec[1] + rvec[2] * rvec[2])
    axis = np.array([rvec[0] / theta, rvec[1] / theta, rvec[2] / theta], dtype=np.float64)
    half = theta * 0.5
    s = np.sin(half)
    q = np.array([axis[0] * s, axis[1] * s, axis[2] * s, np.cos(half)])
    q = q / np.sqrt(np.sum(q ** 2))                          # create_from_axis_rotation normalises ...
    return q / np.sqrt(np.sum(q ** 2))                       # ... and convert_rvec_to_quaternion normalises again


def rotation_from_quaternion(q):
    """pyrr `Quaternion.matrix33` (matrix33.create_from_quaternion) as used at geometric_vision.py:291, :189."""
    qx, qy, qz, qw = [float(v) for v in q]
    sqw, sqx, sqy, sqz = qw * qw, qx * qx, qy * qy, qz * qz
    inv = 1.0 / (sqx + sqy + sqz + sqw)
    return np.array([
        [(sqx - sqy - sqz + sqw) * inv, 2.0 * (qx * qy - qz * qw) * inv, 2.0 * (qx * qz + qy * qw) * inv],
        [2.0 * (qx * qy + qz * qw) * inv, (-sqx + sqy - sqz + sqw) * inv, 2.0 * (qy * qz - qx * qw) * inv],
        [2.0 * (qx * qz - qy * qw) * inv, 2.0 * (qy * qz + qx * qw) * inv, (-sqx - sqy + sqz + sqw) * inv]])


def _rotation_from_rvec(rvec):
    return rotation_from_quaternion(quaternion_from_rvec(rvec))


def solve_pnp_quat(canonical_points, projections, camera_K):
    """geometric_vision.py:43-116: cv2 EPnP, then ITERATIVE refinement from that guess.
    -> (ok, translation [3], quaternion xyzw [4]) -- the pose the reference reports (analysis.py:808-880)."""
    import cv2
    pts = np.asarray(canonical_points, np.float64)
    prj = np.asarray(projections, np.float64)
    if len(pts) != len(prj):
        raise AssertionError("Expected canonical_points and projections to have the same length, but they are "
                             "length {} and {}.".format(len(pts), len(prj)))          # geometric_vision.py:54-58
    if len(pts) == 0:
        return False, None, None
    try:
        ok, rvec, tvec = cv2.solvePnP(pts.reshape(-1, 1, 3), prj.reshape(-1, 1, 2), camera_K, np.array([]),
                                      flags=cv2.SOLVEPNP_EPNP)
        ok, rvec, tvec = cv2.solvePnP(pts.reshape(-1, 1, 3), prj.reshape(-1, 1, 2), camera_K, np.array([]),
                                      flags=cv2.SOLVEPNP_ITERATIVE, useExtrinsicGuess=True, rvec=rvec, tvec=tvec)
        return bool(ok), tvec[:, 0], quaternion_from_rvec(rvec[:, 0])
    except Exception:                       # the reference swallows solver failures the same way (:111-114)
        return False, None, None


def solve_pnp(canonical_points, projections, camera_K):
    """-> (ok, translation [3], rotation [3,3]); see solve_pnp_quat."""
    ok, t, q = solve_pnp_quat(canonical_points, projections, camera_K)
    return (ok, t, rotation_from_quaternion(q)) if ok else (False, None, None)


def post_process_batch(scores, cts_wreg, trans_inv, out_thresh):
    """post_process (sgta_detector.py:929-942 -> post_process.py:93-117), merge_outputs (:955-961) and
    _get_final_kps (:608-651, is_ct branch) for a whole batch: the best-scoring detection of every class, in
    raw-image pixels; MISSING where nothing passed `out_thresh`.  scores [B,K], cts_wreg [B,K,2] float32;
    trans_inv: the float32 inverse output affine (post_process.py:102-103)."""
    B, K = scores.shape
    ones = np.ones((B * K, 3), np.float32)
    ones[:, :2] = cts_wreg.reshape(-1, 2)
    raw = np.dot(trans_inv, ones.transpose()).transpose()[:, :2].reshape(B, K, 2)   # image.py:20-26
    keep = (scores >= out_thresh) & (scores > out_thresh)
    out = np.full((B, K, 2), MISSING)
    out[keep] = raw[keep]
    return out


def post_process_device(scores, cts_wreg, trans_inv, out_thresh, out=None):
    """`post_process_batch` on the device (sgta_post_process): scores [B,K] / cts_wreg [B,K,2] float32 CUDA tensors ->
    kps_raw [B,K,2] float64 CUDA tensor (MISSING where nothing passed `out_thresh`)."""
    import ctypes
    from . import _lib
    B, K = scores.shape
    scores, cts_wreg = scores.contiguous().float(), cts_wreg.contiguous().float()
    if out is None:
        out = torch.empty(B, K, 2, device=scores.device, dtype=torch.float64)
    t6 = (ctypes.c_float * 6)(*[float(v) for v in np.asarray(trans_inv, np.float32).reshape(6)])
    _lib.call("sgta_post_process", _lib.ptr(scores), _lib.ptr(cts_wreg), _lib.ptr(out), t6, float(out_thresh),
              float(MISSING), B, K, _lib.stream())
    return out


def is_pnp_pose(prev_pos, prev_projs, next_pos, prev_projs_all, camera_K):
    """`is_pnp` plus the pose it solved: -> (prev_kp_projs, next_kp_projs_est, pose [7] = t xyz + quaternion xyzw, or
    None when PnP failed).  The pose of frame f-1's detections is what the sequence runner reports per frame."""
    ok, t, q = solve_pnp_quat(prev_pos, prev_projs, camera_K)
    if not ok:
        return prev_projs_all, prev_projs_all, None
    T = np.eye(4)
    T[:3, :3] = rotation_from_quaternion(q)
    T[:3, -1] = t
    hom = np.hstack((next_pos, np.ones((next_pos.shape[0], 1))))
    aligned = np.transpose(np.matmul(T, np.transpose(hom)))[:, :3]
    est = np.transpose(np.matmul(camera_K, np.transpose(aligned)))
    est[:, 0] /= est[:, 2]
    est[:, 1] /= est[:, 2]
    return prev_projs_all, est[:, :2], np.concatenate([np.asarray(t, np.float64), np.asarray(q, np.float64)])


def is_pnp(prev_pos, prev_projs, next_pos, prev_projs_all, camera_K):
    """geometric_vision.py:283-310: pose of the previous frame's detections, applied to the current
    frame's keypoint positions and projected.  -> (prev_kp_projs, next_kp_projs_est)"""
    return is_pnp_pose(prev_pos, prev_projs, next_pos, prev_projs_all, camera_K)[:2]


class LockstepDetector:
    def __init__(self, engine, opt=None, camera_K=None, raw_size=(640, 360), workers=8):
        self.eng = engine
        self.opt = opt if opt is not None else engine.opt
        self.K = np.array(DEFAULT_K if camera_K is None else camera_K, dtype=np.float64)
        self.raw_w, self.raw_h = raw_size
        self.B, self.S = engine.B, engine.S
        self.q = self.S // int(getattr(self.opt, "down_ratio", 4))
        self.n_kp = int(getattr(self.opt, "num_classes", 7))
        self.out_thresh = max(float(getattr(self.opt, "track_thresh", 0.001)), float(getattr(self.opt, "out_thresh", -1)))
        # fix_res pre-processing geometry (sgta_detector.py:354-357, :375-380)
        c = np.array([self.raw_w / 2.0, self.raw_h / 2.0], dtype=np.float32)
        s = max(self.raw_h, self.raw_w) * 1.0
        self.trans_input = PR.get_affine_transform(c, s, 0, [self.S, self.S])
        self.trans_output = PR.get_affine_transform(c, s, 0, [self.q, self.q])
        # post_process: inverse output affine in float32 (post_process.py:102-103)
        self.trans_inv = PR.get_affine_transform(c, s, 0, (self.q, self.q), inv=1).astype(np.float32)
        self.pool = cf.ThreadPoolExecutor(max_workers=workers) if workers > 1 else None
        dev = engine.dev
        self._c_in = torch.zeros(2, self.B, self.n_kp, 2, dtype=torch.float64).pin_memory()
        self._c_out = torch.zeros(2, self.B, self.n_kp, 2, dtype=torch.float64).pin_memory()
        self._d_in = torch.zeros(2, self.B, self.n_kp, 2, dtype=torch.float64, device=dev)
        self._d_out = torch.zeros(2, self.B, self.n_kp, 2, dtype=torch.float64, device=dev)
        self._res = torch.zeros(self.B, self.n_kp, 2, dtype=torch.float64).pin_memory()     # kps_raw
        self._res_sc = torch.zeros(self.B, self.n_kp, dtype=torch.float32).pin_memory()     # scores
        self._d_res = torch.zeros(self.B, self.n_kp, 2, dtype=torch.float64, device=dev)
        self._done = torch.cuda.Event()
        self._begun = False
        self.reset()

    def reset(self):
        self.frame = 0
        self.detected_kps = np.full((self.B, self.n_kp, 2), MISSING)
        self.last_poses = np.full((self.B, 7), np.nan)       # PnP pose (t xyz, q xyzw) of the PREVIOUS frame's detections
        self.timing = {"host_pnp": 0.0, "host_post": 0.0, "steps": 0}

    # ------------------------------------------------------------------ host side of one clip
    def _clip_centres(self, b, x3d_prev, x3d_next):
        """sgta_detector.py:501-547 up to the four rendering calls: raw-pixel keypoints of the two
        prior maps of clip b, or None when nothing was detected (all-zero maps, :523-526)."""
        kps = self.detected_kps[b]
        good = np.unique(np.where(kps > MISSING)[0])
        if len(good) == 0:
            return None, None, None
        return is_pnp_pose(x3d_prev[good], kps[good], x3d_next, kps, self.K)

    def solve_poses(self, x3d):
        """PnP pose [B,7] (t xyz + quaternion xyzw; NaN = nothing detected / solver failed) of the CURRENT
        `detected_kps` against keypoint positions x3d [B,n_kp,3]: the per-frame pose the reference computes offline
        from the saved detections (analysis.py:808-880 -> geometric_vision.solve_pnp).  During a sequence the same
        solve happens inside `begin` of the next frame (`last_poses`); this is for the final frame."""
        def one(b):
            kps = self.detected_kps[b]
            good = np.unique(np.where(kps > MISSING)[0])
            if len(good) == 0:
                return None
            ok, t, q = solve_pnp_quat(x3d[b][good], kps[good], self.K)
            return np.concatenate([t, q]) if ok else None
        res = list(self.pool.map(one, range(self.B))) if self.pool else [one(b) for b in range(self.B)]
        return np.stack([np.full(7, np.nan) if r is None else r for r in res])

    # ------------------------------------------------------------------ one frame of every clip
    def step(self, images, x3d_prev=None, x3d_next=None):
        """images [B,3,S,S] fp32 (device or pinned host), or raw frames [B,raw_h,raw_w,3] uint8 (numpy / torch,
        host or device) that are pre-processed on the device.  x3d_prev / x3d_next: [B,n_kp,3] keypoint
        positions w.r.t. the camera in the previous / this frame (ignored at frame 0).
        Returns {'kps_raw' [B,n_kp,2] float64 (MISSING = not detected), 'scores' [B,n_kp]}."""
        self.begin(images, x3d_prev, x3d_next)
        return self.finish()

    def begin(self, images, x3d_prev=None, x3d_next=None):
        """First half of `step`: the host PnP of every clip, then everything the device has to do for this
        frame is ENQUEUED (uploads, pre-processing, prior maps, network, decode, result download) and the
        call returns.  `finish` waits for it.  Between the two the host is free -- `ClipGroups` uses that to
        run another group's PnP under this group's device work."""
        eng, inp = self.eng, self.eng.inp
        t0 = time.perf_counter()
        if isinstance(images, np.ndarray):
            images = torch.from_numpy(images)
        raw = images.dtype == torch.uint8

        def load_x():
            if raw:                                              # pre_process (:368-399) on the device
                if tuple(images.shape[1:]) != (self.raw_h, self.raw_w, 3):
                    raise ValueError("raw frames must be [B,%d,%d,3] uint8" % (self.raw_h, self.raw_w))
                PP.warp_normalize(images.to(inp["x"].device, non_blocking=True), self.trans_input,
                                  (self.S, self.S), out=inp["x"])
            else:
                inp["x"].copy_(images, non_blocking=True)

        if self.frame == 0:
            # _get_additional_inputs (:415-454): all-zero priors, pre_images = images (:157-159)
            for k in ("pre_hm", "repro_hm", "pre_hm_cls", "repro_hm_cls"):
                inp[k].zero_()
            load_x()
            inp["pre_img"].copy_(inp["x"])
        else:
            inp["pre_img"].copy_(inp["x"])                       # self.pre_images = images (:204)
            load_x()
            jobs = [(b, x3d_prev[b], x3d_next[b]) for b in range(self.B)]
            res = list(self.pool.map(lambda a: self._clip_centres(*a), jobs)) if self.pool else \
                [self._clip_centres(*a) for a in jobs]
            far = np.full((self.n_kp, 2), -1.0)                  # outside the raw image -> (0,0) -> nothing drawn
            prev = np.stack([far if r[0] is None else r[0] for r in res])
            nxt = np.stack([far if r[1] is None else r[1] for r in res])
            self.last_poses = np.stack([np.full(7, np.nan) if r[2] is None else r[2] for r in res])
            for i, pts in enumerate((prev, nxt)):
                self._c_in[i].copy_(torch.from_numpy(PR.affine_transform_and_clip(
                    pts, self.trans_input, self.S, self.S, self.raw_w, self.raw_h)))
                self._c_out[i].copy_(torch.from_numpy(PR.affine_transform_and_clip(
                    pts, self.trans_output, self.q, self.q, self.raw_w, self.raw_h)))
            self._d_in.copy_(self._c_in, non_blocking=True)
            self._d_out.copy_(self._c_out, non_blocking=True)
            PR.render_priors(self._d_in[0], self._d_out[0], self.S, self.q, hm=inp["pre_hm"], hm_cls=inp["pre_hm_cls"])
            PR.render_priors(self._d_in[1], self._d_out[1], self.S, self.q, hm=inp["repro_hm"], hm_cls=inp["repro_hm_cls"])
        t1 = time.perf_counter()
        dets = eng.infer()
        # post_process + merge_outputs + _get_final_kps on the device: raw-image keypoints (float64) and scores come
        # back in two small copies
        post_process_device(dets["scores"].view(self.B, self.n_kp), dets["cts_wreg"].view(self.B, self.n_kp, 2),
                            self.trans_inv, self.out_thresh, out=self._d_res)
        self._res.copy_(self._d_res, non_blocking=True)
        self._res_sc.copy_(dets["scores"].view(self.B, self.n_kp), non_blocking=True)
        self._done.record(torch.cuda.current_stream(eng.dev))
        self.timing["host_pnp"] += t1 - t0
        self._begun = True

    def finish(self):
        """Second half of `step`: wait for the frame enqueued by `begin`, post-process on the host."""
        if not self._begun:
            raise RuntimeError("finish() without begin()")
        self._begun = False
        self._done.synchronize()
        t2 = time.perf_counter()
        scores = self._res_sc.numpy().copy()
        self.detected_kps = self._res.numpy().copy()
        t3 = time.perf_counter()
        self.timing["host_post"] += t3 - t2
        self.timing["steps"] += 1
        self.frame += 1
        return {"kps_raw": self.detected_kps.copy(), "scores": scores}


class ClipGroups:
    """G lock-step groups of clips, one engine each, run SKEWED: while the device works on group g's frame,
    the host solves the PnP of group g+1 (and post-processes group g-1).  Clips are independent
    (inference.py:201-205 builds a fresh detector per clip), so splitting the batch changes nothing that is
    computed -- only the host work (north_star keeps LM/PnP on the host) stops serialising with the device:

        begin(A,f) begin(B,f) | finish(A,f) begin(A,f+1) | finish(B,f) begin(B,f+1) | ...

    `run` drives that schedule for a whole sequence."""

    def __init__(self, detectors):
        self.dets = list(detectors)
        self.B = sum(d.B for d in self.dets)
        self.offsets = np.cumsum([0] + [d.B for d in self.dets])

    def reset(self):
        for d in self.dets:
            d.reset()

    def _slice(self, a, g):
        return None if a is None else a[self.offsets[g]:self.offsets[g + 1]]

    def run(self, n_frames, images_fn, x3d_fn=None, before_begin=None):
        """images_fn(f) -> frames of ALL clips for frame f ([B,...] uint8 raw or float32 network inputs);
        x3d_fn(f) -> [B,n_kp,3] keypoint positions w.r.t. the camera in frame f.
        before_begin(g, f, det): optional hook called right before group g's begin of frame f (tests and the
        bench plant detections there).  Returns per-frame {'kps_raw' [B,n_kp,2], 'scores' [B,n_kp], 'pose' [B,7] = PnP
        pose (t xyz + quaternion xyzw) of that frame's detections against x3d_fn(f), NaN where there is none}."""
        G = len(self.dets)
        out = [dict() for _ in range(n_frames)]
        self.reset()                                   # a reused detector starts its clips again at frame 0

        def begin(g, f):
            d = self.dets[g]
            if before_begin is not None:
                before_begin(g, f, d)
            prev = self._slice(x3d_fn(f - 1), g) if (x3d_fn is not None and f > 0) else None
            nxt = self._slice(x3d_fn(f), g) if (x3d_fn is not None and f > 0) else None
            d.begin(self._slice(images_fn(f), g), prev, nxt)

        def finish(g, f):
            out[f][g] = self.dets[g].finish()

        def pose_of(g, f, final):
            """pose of frame f's detections: solved inside begin(g, f + 1), or explicitly after the last frame"""
            d = self.dets[g]
            if x3d_fn is None or not hasattr(d, "last_poses"):
                return np.full((d.B, 7), np.nan)
            return d.solve_poses(self._slice(x3d_fn(f), g)) if final else d.last_poses.copy()

        for g in range(G):
            begin(g, 0)
        for f in range(n_frames):
            for g in range(G):
                finish(g, f)
                if f + 1 < n_frames:
                    begin(g, f + 1)
                    out[f][g]["pose"] = pose_of(g, f, False)
                else:
                    out[f][g]["pose"] = pose_of(g, f, True)
        return [{k: np.concatenate([fr[g][k] for g in range(G)]) for k in ("kps_raw", "scores", "pose")} for fr in out]
