"""CPU tests of the multi-GPU sequence runner (sgtapose_b200/runner.py) with stand-in detectors: clip sharding,
waves (ragged last wave padded), per-frame pose hand-over from `begin` of the next frame / `solve_poses` of the
last one, and the final all-gather over gloo with world_size 2 (the N > 1 path of bench.py's `pipeline` leg;
reference loop: sgtapose/inference.py:186-294)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


class _Det:
    """Lock-step detector stand-in: 'detects' keypoints that encode (clip id, frame), 'solves' poses that encode
    the x3d it was given."""
    n_kp = 7

    def __init__(self, B):
        self.B = B
        self.reset()

    def reset(self):
        self.frame = 0
        self.last_poses = np.full((self.B, 7), np.nan)
        self._cur = None

    def begin(self, images, x3d_prev=None, x3d_next=None):
        assert images.shape == (self.B, 2)
        if self.frame == 0:
            assert x3d_prev is None
        else:
            # pose of the previous frame's detections: depends on that frame's x3d and detections
            self.last_poses = self._pose(x3d_prev, self.kps)
        self._cur = images

    def finish(self):
        clip, f = self._cur[:, 0], self._cur[:, 1]
        assert np.all(f == self.frame)
        self.kps = (clip[:, None, None] * 1000 + f[:, None, None] * 10 + np.arange(14).reshape(1, 7, 2)).astype(np.float64)
        self.frame += 1
        return {"kps_raw": self.kps.copy(), "scores": np.tile(clip[:, None] + 0.5, (1, 7)).astype(np.float32)}

    @staticmethod
    def _pose(x3d, kps):
        return np.concatenate([x3d[:, 0, :], kps[:, :4, 0]], axis=1)

    def solve_poses(self, x3d):
        return self._pose(x3d, self.kps)


def _images(ids, f):
    return np.stack([np.array([c, f], dtype=np.float64) for c in ids])


def _x3d(ids, f):
    return np.stack([np.full((7, 3), 100.0 * c + f) for c in ids])


def _expected(n_clips, n_frames):
    kps = np.zeros((n_clips, n_frames, 7, 2))
    pose = np.zeros((n_clips, n_frames, 7))
    for c in range(n_clips):
        for f in range(n_frames):
            kps[c, f] = c * 1000 + f * 10 + np.arange(14).reshape(7, 2)
            pose[c, f, :3] = 100.0 * c + f
            pose[c, f, 3:] = kps[c, f, :4, 0]
    return kps, pose


def test_runner_single_rank_waves_and_poses():
    from sgtapose_b200.runner import SequenceRunner
    n_clips, n_frames = 7, 4                              # waves of 2 + 1 = 3 clips -> 3 waves, the last one ragged
    hooks = []
    r = SequenceRunner([_Det(2), _Det(1)])
    out = r.run(n_clips, n_frames, _images, _x3d, before_begin=lambda ids, f, d: hooks.append((tuple(ids), f)))
    kps, pose = _expected(n_clips, n_frames)
    assert np.array_equal(out["kps_raw"], kps)
    assert np.array_equal(out["pose"], pose)
    assert np.array_equal(out["scores"][:, 0, 0], np.arange(n_clips) + 0.5)
    assert r.timing["waves"] == 3 and r.timing["local_clips"] == 7
    assert hooks[:2] == [((0, 1), 0), ((2,), 0)] and ((6, 6), 0) in hooks    # padded wave repeats its last clip


def _worker(rank, world, port, n_clips, n_frames, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from sgtapose_b200.runner import SequenceRunner
        r = SequenceRunner([_Det(2), _Det(2)], world=world, rank=rank)
        out = r.run(n_clips, n_frames, _images, _x3d)
        q.put((rank, out["kps_raw"], out["pose"], out["scores"], r.timing["local_clips"]))
    finally:
        dist.destroy_process_group()


def test_runner_gloo_world2_matches_single_rank():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    n_clips, n_frames, world = 9, 3, 2                    # 5 + 4 clips; rank 0: waves of 4 + 1 (ragged)
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_clips, n_frames, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    kps, pose = _expected(n_clips, n_frames)
    assert sorted(r[4] for r in res) == [4, 5]
    for rank, k, p_, s_, _ in res:
        assert np.array_equal(k, kps), rank
        assert np.array_equal(p_, pose), rank
        assert np.array_equal(s_[:, 0, 0], np.arange(n_clips) + 0.5), rank


def test_runner_edge_cases_empty_shard_and_single_frame():
    """No clips at all (an idle rank keeps its place in the collective with an empty shard) and one-frame clips
    (every pose comes from `solve_poses`, none from a next frame's `begin`)."""
    from sgtapose_b200.runner import SequenceRunner
    r = SequenceRunner([_Det(2), _Det(2)])
    out = r.run(0, 3, _images, _x3d)
    assert out["kps_raw"].shape == (0, 3, 7, 2) and out["pose"].shape == (0, 3, 7) and out["scores"].shape == (0, 3, 7)
    assert r.timing["waves"] == 0
    out = r.run(3, 1, _images, _x3d)
    kps, pose = _expected(3, 1)
    assert np.array_equal(out["kps_raw"], kps) and np.array_equal(out["pose"], pose)
    # a rank beyond the clip count owns nothing and must not touch its detectors
    r2 = SequenceRunner([_Det(2)], world=4, rank=3)
    out = r2.run(2, 2, _images, _x3d, gather=False)
    assert out["kps_raw"].shape[0] == 0 and r2.timing["local_clips"] == 0
