"""BASELINE configs[4]: training step (fwd + bwd through DCN and the temporal attention) under DDP.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/train_step_ddp.py \
      [--batch 16] [--size 384] [--steps 5]

One process per GPU, NCCL gradient all-reduce (DistributedDataParallel, find_unused_parameters=True because
base.base_layer never runs, trainer_parallel.py:143-145), MSE on sigmoid(hm) + L1 on reg/tracking like
Loss.forward (trainer_parallel.py:88-127), SGD step.  Module-tree path (networks.py): DCN fwd/bwd and attention
fwd/bwd are this repo's kernels, plain convs are cuDNN.  Prints one JSON line on rank 0 (device-timed, max over ranks).
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--size", type=int, default=384)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    args = ap.parse_args()
    from sgtapose_b200 import config, networks, synth
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    opt = config.default_opt()
    model = networks.create_model(config.ARCH, dict(config.HEADS), dict(config.HEAD_CONV), opt)
    model.load_state_dict(synth.synthetic_state_dict(model.state_dict(), seed=317))
    model = model.to(dev).train()
    net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], find_unused_parameters=True) \
        if world > 1 else model
    optim = torch.optim.SGD(model.parameters(), lr=1e-5)
    ins = [t.to(dev) for t in synth.synthetic_inputs(args.batch, args.size, seed=317 + rank, frame=1)]
    q = args.size // 4
    tgt = {"hm": torch.rand(args.batch, 7, q, q, device=dev), "reg": torch.rand(args.batch, 2, q, q, device=dev),
           "tracking": torch.rand(args.batch, 2, q, q, device=dev)}

    def step():
        out = net(*ins)[0]
        hm = torch.clamp(out["hm"].sigmoid(), 1e-4, 1 - 1e-4)                       # utils.py:15-17
        loss = torch.nn.functional.mse_loss(hm, tgt["hm"]) + \
            torch.nn.functional.l1_loss(out["reg"], tgt["reg"]) + torch.nn.functional.l1_loss(out["tracking"], tgt["tracking"])
        optim.zero_grad(set_to_none=True)
        loss.backward()
        optim.step()
        return loss

    for _ in range(args.warmup):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        ms = t.item() / args.steps
        print(json.dumps({"workload": "configs[4]: training step fwd+bwd, DDP gradient all-reduce", "n_gpus": world,
                          "batch_per_gpu": args.batch, "size": args.size, "ms_per_step": ms,
                          "samples_per_s": args.batch * world / (ms / 1e3), "loss": float(loss), "dtype": "f32"}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
