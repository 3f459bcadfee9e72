"""CPU tests (gloo, world_size 2) of the clip sharding and the result all-gather of the
multi-GPU path (sgtapose_b200/shard.py)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sgtapose_b200 import shard


def test_shard_clips_partition():
    for n, w in [(7, 2), (64, 8), (3, 4), (0, 2)]:
        seen = sorted(i for r in range(w) for i in shard.shard_clips(n, w, r))
        assert seen == list(range(n))
        sizes = [len(shard.shard_clips(n, w, r)) for r in range(w)]
        assert max(sizes) - min(sizes) <= 1
    assert shard.lockstep_batches(list(range(5)), 2) == [[0, 1], [2, 3], [4]]


def _worker(rank, world, port, n_clips, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ids = shard.shard_clips(n_clips, world, rank)
        # a rank's "poses": [n_local, 7, 2] keypoints, value encodes (clip, keypoint, xy)
        local = torch.tensor([[[c * 100 + k * 2 + d for d in range(2)] for k in range(7)] for c in ids],
                             dtype=torch.float32).reshape(len(ids), 7, 2)
        full = shard.gather_results(local, n_clips, world, rank)
        t = torch.tensor([1.0 + rank])
        dist.all_reduce(t, op=dist.ReduceOp.MAX)          # the bench's max-over-ranks timing reduction
        q.put((rank, full, float(t)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_clips", [5, 1])               # ragged 3 + 2; one clip: rank 1 owns nothing
def test_gather_results_gloo_world2(n_clips):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world = 2
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_clips, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = torch.tensor([[[c * 100 + k * 2 + d for d in range(2)] for k in range(7)] for c in range(n_clips)],
                        dtype=torch.float32)
    for rank, full, tmax in res:
        assert torch.equal(full, want), rank
        assert tmax == 2.0
