"""Multi-GPU sequence runner: N clips of F frames -> per-frame keypoints and poses of every clip, on every rank.

The reference runs clips one after the other, a fresh detector per clip and the frames of a clip in order
(sgtapose/inference.py:186-294: `detector = SGTADetector(...)` per video, `detector.run(img, j, json_path, ...)` per
frame -> `detected_kps_np`; the poses are then solved per frame from the saved detections,
analysis.py:808-880 -> geometric_vision.solve_pnp).  Clips are independent and frames serial, so (SURVEY.md 8e,
BASELINE configs[2] and [3]):

  * `shard_clips` gives rank r the clips r, r+R, ... (no data-path collective, weights replicated);
  * a rank cuts its clips into WAVES of `n_groups x group_size` clips; a wave is run frame by frame by
    `ClipGroups` (detector.py): `n_groups` lock-step groups, one engine each, skewed so that the host PnP of one
    group runs under the device work of the other;
  * per clip and frame the rank keeps [kps_raw (n_kp x 2) | scores (n_kp) | pose (t xyz, q xyzw)] and ONE
    `all_gather` at the end (`shard.gather_results`: NCCL on the B200 box, gloo in the CPU tests) gives every rank
    the [n_clips, F, 3*n_kp + 7] result in global clip order -- KBs, latency-bound, nothing to fuse.

A ragged last wave is padded by repeating its last clip (the engines have a fixed lock-step batch); the padding
results are dropped.
"""
import time

import numpy as np
import torch

from . import shard
from .detector import ClipGroups

N_POSE = 7


class SequenceRunner:
    def __init__(self, detectors, world=1, rank=0, device=None):
        """detectors: the lock-step detectors of ONE wave (LockstepDetector-like: B, begin, finish, reset,
        last_poses / solve_poses); they are reused for every wave of this rank."""
        self.groups = ClipGroups(detectors)
        self.wave = self.groups.B
        self.world, self.rank = world, rank
        self.device = device
        self.n_kp = int(getattr(detectors[0], "n_kp", 7))
        self.timing = {}

    def run(self, n_clips, n_frames, images_fn, x3d_fn, before_begin=None, gather=True):
        """images_fn(clip_ids, f) -> frames of those clips at frame f ([len(clip_ids), ...] raw uint8 or float32);
        x3d_fn(clip_ids, f) -> [len(clip_ids), n_kp, 3] keypoint positions w.r.t. the camera;
        before_begin(clip_ids_of_group, f, det): optional hook right before a group's `begin` of frame f.
        Returns {'kps_raw' [n_clips,F,n_kp,2], 'scores' [n_clips,F,n_kp], 'pose' [n_clips,F,7]} (numpy, global clip
        order, identical on every rank) -- or this rank's clips only with gather=False."""
        ids = shard.shard_clips(n_clips, self.world, self.rank)
        width = 3 * self.n_kp + N_POSE
        local = np.zeros((len(ids), n_frames, width))
        t0 = time.perf_counter()
        for w0 in range(0, len(ids), self.wave):
            real = ids[w0:w0 + self.wave]
            wave_ids = real + [real[-1]] * (self.wave - len(real))       # ragged wave: repeat the last clip
            offs = self.groups.offsets
            hook = None
            if before_begin is not None:
                hook = lambda g, f, d: before_begin(wave_ids[offs[g]:offs[g + 1]], f, d)
            frames = self.groups.run(n_frames, lambda f: images_fn(wave_ids, f), lambda f: x3d_fn(wave_ids, f),
                                     before_begin=hook)
            for f, fr in enumerate(frames):
                n = len(real)
                local[w0:w0 + n, f, :2 * self.n_kp] = fr["kps_raw"][:n].reshape(n, -1)
                local[w0:w0 + n, f, 2 * self.n_kp:3 * self.n_kp] = fr["scores"][:n]
                local[w0:w0 + n, f, 3 * self.n_kp:] = fr["pose"][:n]
        t1 = time.perf_counter()
        full = local
        if gather and self.world > 1:
            t = torch.from_numpy(local)
            if self.device is not None:
                t = t.to(self.device)
            full = shard.gather_results(t, n_clips, self.world, self.rank).cpu().numpy()
        t2 = time.perf_counter()
        self.timing = {"run_s": t1 - t0, "gather_s": t2 - t1, "local_clips": len(ids),
                       "waves": (len(ids) + self.wave - 1) // self.wave}
        k = self.n_kp
        return {"kps_raw": full[:, :, :2 * k].reshape(full.shape[0], n_frames, k, 2),
                "scores": full[:, :, 2 * k:3 * k], "pose": full[:, :, 3 * k:]}
