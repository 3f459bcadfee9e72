"""Clip sharding across the GPUs of one box and the one collective of the inference path.

Clips are independent (reference sgtapose/inference.py:201-205 builds a fresh detector per clip)
while frames inside a clip are serial, so rank r owns clips r, r+R, r+2R, ... and runs them as
lock-step batches; weights are replicated.  The only exchange is an all-gather of the decoded
per-frame results at the end -- KBs, latency-bound; `torch.distributed` with the
"nccl" backend on the B200 box and "gloo" in the CPU tests.
"""
import torch
import torch.distributed as dist


def shard_clips(n_clips, world, rank):
    """Indices of the clips rank `rank` owns (round-robin: balanced to within one clip)."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world %d" % (rank, world))
    return list(range(rank, n_clips, world))


def lockstep_batches(clip_ids, batch):
    """Cut a rank's clips into lock-step batches of at most `batch` clips."""
    return [clip_ids[i:i + batch] for i in range(0, len(clip_ids), batch)]


def gather_results(local, n_clips, world, rank):
    """local: [n_local, ...] results of this rank's clips in shard_clips order.  Returns the
    [n_clips, ...] tensor in global clip order on every rank (one all_gather; ragged shards are
    padded to the largest shard)."""
    if world == 1:
        return local
    per = (n_clips + world - 1) // world
    pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    out = torch.empty((n_clips,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    for r in range(world):
        ids = shard_clips(n_clips, world, r)
        out[ids] = parts[r][:len(ids)]
    return out
