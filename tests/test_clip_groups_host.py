"""Host logic of the skewed clip-group schedule (sgtapose_b200/detector.py::ClipGroups) with stand-in detectors:
the order of begin / finish calls, the slices each group receives and the assembly of per-frame results.  CPU test."""
import numpy as np


class _FakeDet:
    def __init__(self, B, log, name):
        self.B, self.log, self.name, self.frame = B, log, name, 0
        self.pending = None

    def reset(self):
        self.frame = 0

    def begin(self, images, x3d_prev=None, x3d_next=None):
        assert self.pending is None, "begin() twice without finish()"
        assert images.shape[0] == self.B
        if self.frame == 0:
            assert x3d_prev is None and x3d_next is None
        else:
            assert x3d_prev.shape == x3d_next.shape == (self.B, 7, 3)
        self.log.append(("begin", self.name, self.frame))
        self.pending = (images.copy(), None if x3d_next is None else x3d_next.copy())

    def finish(self):
        images, x3d = self.pending
        self.pending = None
        self.log.append(("finish", self.name, self.frame))
        self.frame += 1
        # result = something traceable to the inputs of THIS frame and THIS group's clips
        kps = np.repeat(images.reshape(self.B, -1)[:, :1, None], 7, 1).repeat(2, 2).astype(np.float64)
        return {"kps_raw": kps, "scores": images.reshape(self.B, -1)[:, :7].astype(np.float32)}


def test_skewed_schedule_order_slices_and_assembly():
    from sgtapose_b200.detector import ClipGroups
    log = []
    groups = ClipGroups([_FakeDet(3, log, "A"), _FakeDet(2, log, "B")])
    assert groups.B == 5 and groups.offsets.tolist() == [0, 3, 5]
    n_frames = 3
    frames = [np.arange(5 * 8, dtype=np.float32).reshape(5, 8) + 100 * f for f in range(n_frames)]
    x3d = [np.full((5, 7, 3), float(f)) for f in range(n_frames)]
    hooks = []
    out = groups.run(n_frames, lambda f: frames[f], lambda f: x3d[f], before_begin=lambda g, f, d: hooks.append((g, f)))
    # begin(A,0) begin(B,0) | finish(A,0) begin(A,1) finish(B,0) begin(B,1) | ... | finish(A,2) finish(B,2)
    want = [("begin", "A", 0), ("begin", "B", 0)]
    for f in range(n_frames):
        for g in "AB":
            want.append(("finish", g, f))
            if f + 1 < n_frames:
                want.append(("begin", g, f + 1))
    assert log == want
    assert hooks == [(0, 0), (1, 0), (0, 1), (1, 1), (0, 2), (1, 2)]
    for f in range(n_frames):
        assert out[f]["scores"].shape == (5, 7) and out[f]["kps_raw"].shape == (5, 7, 2)
        assert np.array_equal(out[f]["scores"], frames[f][:, :7])           # group slices re-assembled in clip order
        assert np.array_equal(out[f]["kps_raw"][:, 0, 0], frames[f][:, 0].astype(np.float64))
