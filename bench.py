#!/usr/bin/env python
"""bench.py -- pose frames/s of the SGTAPose per-frame dense inference path on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]

One "step" = one pass of the hot path over one lock-step batch of B clips (one frame-pair
each, 384x384, 7 Panda keypoints): network forward (DLA-34 x2 + structure-prior attention +
DLAUp/IDAUp with 16 DCNs + heads) -> sigmoid -> live heatmap decode.  Workload =
BASELINE.json configs[1] ("same model fp32 batch 32 on 1xB200"); N>1 shards clips across
ranks (weak scaling, B clips per GPU) and all-gathers the decoded keypoints over NCCL.

Prints ONE JSON line (rank 0).  `value` = device-timed frames/s with inputs resident in
HBM; `e2e` = the same through the public API with HOST (pinned) inputs, H2D and D2H inside
the timed region; `roofline` = the DCN kernels (tensor bound) timed live with CUDA events;
`cpu_baseline` = the oracle port (torch CPU + torchvision deform_conv2d + numpy decode) on
the host cores, bounded sample.  `--impl reference` times that CPU path alone.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

S = 384
DCN_GFLOP_PER_FRAME = 7.984          # sum 2*Cout*9Cin*H*W over the 16 DCNs
# dram__bytes_read.sum + dram__bytes_write.sum summed over the 16 DCN launches of ONE step at the bench config
# (fp32 mode, 32 clips): one `ncu --set full` capture, profiles/r1_ncu_full_convs_fp32_b32.txt (999.0 + 214.1 MB)
DCN_DRAM_BYTES_PER_STEP_FP32_B32 = 1213.1e6
TOTAL_GFLOP_PER_FRAME = 56.6


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i] == "Active"})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx[0] if mx else None,
                "reasons": reasons, "samples": len(sm)}


def build_model(device):
    from sgtapose_b200 import config, networks, synth
    opt = config.default_opt()
    model = networks.create_model(config.ARCH, dict(config.HEADS), dict(config.HEAD_CONV), opt)
    sd = synth.synthetic_state_dict(model.state_dict(), seed=317)
    model.load_state_dict(sd)
    return model.eval().to(device), sd, opt


def cpu_reference_fps(sd, steps, warmup, B=1):
    """The reference's CPU path restated (oracle/): fp32 torch + torchvision deform_conv2d +
    numpy live decode, all host threads."""
    from oracle import decode as odec
    from oracle import model as omodel
    from sgtapose_b200 import synth
    torch.set_num_threads(os.cpu_count())
    ins = synth.synthetic_inputs(B, S, seed=317, frame=1)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        out = omodel.forward(sd, *ins)[0]
        hm = torch.sigmoid(out["hm"]).numpy()
        odec.dream_generic_decode(hm, out["reg"].numpy(), out["tracking"].numpy())
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    times.sort()
    med = times[len(times) // 2]
    return B / med, med, os.cpu_count()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from sgtapose_b200 import config, networks, synth
    model = networks.create_model(config.ARCH, dict(config.HEADS), dict(config.HEAD_CONV), config.default_opt())
    sd = synth.synthetic_state_dict(model.state_dict(), seed=317)
    fps, med, cores = cpu_reference_fps(sd, args.steps, args.warmup)
    sample = "batch 1 frame-pair 384x384 per step (the GPU arm runs %d per step), median of %d steps" % (
        args.batch, args.steps)
    line = {"impl": "reference", "metric": "pose_frames_per_sec", "value": fps, "unit": "frames/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": med * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": "configs[1]: DLA-34+DCN SGTAPose forward + live decode, 384x384, "
                                   "2-frame synthetic Panda, fp32", "batch_per_step": 1},
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch.distributed as dist
    from sgtapose_b200 import _lib, decode, synth
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.benchmark = True
    B = args.batch
    if args.dbg:
        _lib.load().sgta_debug_flags(args.dbg)
    host = [t.pin_memory() for t in synth.synthetic_inputs(B, S, seed=317 + rank, frame=1)]
    resident = [t.to(dev, non_blocking=True) for t in host]
    h2d = sum(t.numel() * t.element_size() for t in host)
    if args.engine == "eager":
        model, sd, opt = build_model(dev)
        model.skip_dead_levels = bool(args.skip_dead_levels)

        def step(inputs):
            with torch.no_grad():
                out = model(*inputs)[0]
                out["hm"] = out["hm"].sigmoid_()                    # sgta_detector.py:854-862
                return decode.dream_generic_decode(out, K=7, opt=opt)
        eager_pass = lambda: step(resident)
    else:
        from sgtapose_b200 import config, engine, networks
        opt = config.default_opt()
        tmpl = networks.create_model(config.ARCH, dict(config.HEADS), dict(config.HEAD_CONV), opt)
        sd = synth.synthetic_state_dict(tmpl.state_dict(), seed=317)
        del tmpl
        eng = engine.InferenceEngine(sd, opt, batch=B, size=S, mode=args.mode, device=dev, fuse_sigmoid=True,
                                     skip_dead_levels=bool(args.skip_dead_levels), use_graph=True)

        def step(inputs):
            return eng.infer(*inputs)

        def eager_pass():
            with torch.no_grad():
                eng._run()

    def step_e2e():
        # eager modules take device tensors; the engine copies pinned host inputs straight into
        # its static buffers (H2D inside the timed region either way)
        dets = step([t.to(dev, non_blocking=True) for t in host]) if args.engine == "eager" else step(host)
        res = {k: dets[k].cpu() for k in ("scores", "cts_wreg", "xs", "ys", "tracking")}
        return res

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        step(resident)
    barrier()
    if args.profile_pass:
        # one un-graphed pass between cudaProfilerStart/Stop (ncu --profile-from-start off)
        torch.cuda.profiler.start()
        eager_pass()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return
    l0 = _lib.launch_count()
    eager_pass()                                   # count our kernels in one un-graphed pass
    step_launches = _lib.launch_count() - l0 + (1 if args.engine != "eager" else 0)   # + decode
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        dets = step(resident)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = step_launches * args.steps
    clocks = sampler.stop() if rank == 0 else None

    # end-to-end: host inputs, H2D + D2H inside the timed region.  Engine: the pipelined host API
    # (submit / launch / collect): step i+1's H2D runs on a copy stream while step i computes; every step
    # still copies its own 167.5 MB of inputs from pinned host memory and reads its own results back.
    pipelined = args.engine != "eager" and not args.e2e_sync
    if pipelined:
        def run_e2e(n):
            eng.submit(*host)
            for i in range(n):
                eng.launch()
                if i + 1 < n:
                    eng.submit(*host)
                r = eng.collect()
            return r
        run_e2e(2)
        barrier()
        e0.record()
        res = run_e2e(args.steps)
        e1.record()
        barrier()
        d2h = sum(v.size * v.itemsize for v in res.values())
    else:
        for _ in range(2):
            step_e2e()
        barrier()
        e0.record()
        for _ in range(args.steps):
            res = step_e2e()
        e1.record()
        barrier()
        d2h = sum(v.numel() * v.element_size() for v in res.values())
    ms_e2e = e0.elapsed_time(e1)

    # per-kernel timing of the DCN launches (CUDA events on the launching stream)
    dcn_ms = time_dcn_kernels(eager_pass)
    extra = {}
    if rank == 0 and world == 1 and args.engine != "eager" and not args.no_extras:
        extra["roofline_decode"] = decode_roofline(dev)
        extra["roofline_preprocess"] = preprocess_roofline(dev)
        extra["pipeline"] = clip_pipeline(eng, dev, args.clip_frames, sd)

    t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        gathered = [torch.empty_like(dets["cts_wreg"]) for _ in range(world)]
        dist.all_gather(gathered, dets["cts_wreg"].contiguous())        # per-rank poses -> everyone
    ms, ms_e2e = t.tolist()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks, which = _peaks()
    frames = B * world * args.steps
    value = frames / (ms / 1e3)
    dcn_tflops = DCN_GFLOP_PER_FRAME * B / (dcn_ms / 1e3) / 1e3
    peak_tf = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])
    line = {
        "metric": "pose_frames_per_sec", "value": value, "unit": "frames/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32" if args.mode == "fp32" else "bf16", "data": "synthetic",
        "config": {"workload": "configs[1]: DLA-34+DCN SGTAPose forward + live decode, 384x384, 2-frame "
                               "synthetic Panda, " + args.mode, "batch_per_gpu": B, "clips_sharded_by": "rank",
                   "l2": "inputs+activations per step (>1 GB) exceed the 126 MB L2",
                   "skip_dead_levels": bool(args.skip_dead_levels),
                   "engine": "eager modules + libsgta_b200" if args.engine == "eager" else
                             "InferenceEngine (NHWC, tcgen05 convs, CUDA graph)",
                   "arithmetic": "fp32 activations as fp16 hi+lo planes, 3 tensor-core MMAs per K step, rotating fp32 accumulators"
                                 if args.mode == "fp32" else "bf16 activations and MMAs, fp32 accumulate"},
        "e2e": {"value": frames / (ms_e2e / 1e3), "unit": "frames/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h,
                "api": "InferenceEngine.submit/launch/collect (H2D of step i+1 overlaps step i)" if pipelined
                       else "synchronous infer() per step"},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": "dcn (16 launches/step)", "achieved": dcn_tflops,
                     "peak": peak_tf, "unit": "TFLOP/s", "frac": dcn_tflops / peak_tf,
                     "traffic": DCN_DRAM_BYTES_PER_STEP_FP32_B32 if (args.mode == "fp32" and B == 32 and args.engine != "eager") else None,
                     "traffic_note": "ncu dram bytes, sum over the 16 launches of one step (achieved is also per step)",
                     "peak_source": which + " bf16 sustained", "ms_per_step": dcn_ms},
    }
    line.update(extra)
    if world == 1 and not args.no_cpu_baseline:
        fps, med, cores = cpu_reference_fps(sd, 5, 2)
        line["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                                "sample": "5 timed frame-pairs (batch 1) after 2 warm-ups, median"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def decode_roofline(dev, B=1024):
    """Live heat-map decode alone (BASELINE metric: decode HBM GB/s): B frames of [7,96,96] heat maps per
    launch, algorithmic bytes = one read of hm = 258,048 B/frame, CUDA events, median of 5."""
    from sgtapose_b200 import decode, synth
    hm, _ = synth.synthetic_heatmaps(64, seed=317)
    hm = hm.to(dev).repeat(B // 64, 1, 1, 1).contiguous()
    reg = torch.rand(B, 2, 96, 96, device=dev)
    trk = torch.rand(B, 2, 96, 96, device=dev)
    def med5(**kw):
        for _ in range(3):
            decode.peaks_decode(hm, reg, trk, **kw)
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            decode.peaks_decode(hm, reg, trk, **kw)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return sorted(ts)[2]

    decode.recheck_count(reset=True)
    ms = med5()
    rechecks = decode.recheck_count(reset=True) / 8.0          # 3 warm-up + 5 timed launches
    ms64 = med5(exact64=True)
    peaks, which = _peaks()
    gbs = hm.numel() * 4 / (ms / 1e3) / 1e9
    return {"bound": "hbm", "kernel": "decode_peaks_f32_kernel", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
            "frac": gbs / peaks["hbm_gbs"], "traffic": None, "frames_per_launch": B, "ms": ms,
            "frames_per_s": B / (ms / 1e3), "peak_source": which, "ms_all_float64_kernel": ms64,
            "float64_rechecked_pixels_per_launch": rechecks,
            "note": "float32 blur with a rigorous rounding band; pixels inside the band are re-evaluated with the "
                    "bit-exact float64 scipy restatement (the all-float64 kernel is timed beside it)"}


def preprocess_roofline(dev, B=256):
    """Device pre-processing alone: B raw 640x360 uint8 frames -> [B,3,384,384] fp32 per launch;
    algorithmic bytes = raw frame read once + network input written once; CUDA events, median of 5."""
    import numpy as np
    from sgtapose_b200 import config, preprocess
    rng = np.random.default_rng(317)
    frames = torch.from_numpy(rng.integers(0, 256, (B, 360, 640, 3), dtype=np.uint8)).to(dev)
    opt = config.default_opt(fix_res=True, fix_short=-1, input_h=S, input_w=S, down_ratio=4)
    meta = preprocess.transform_meta(360, 640, opt)
    out = torch.empty(B, 3, S, S, device=dev)
    for _ in range(3):
        preprocess.warp_normalize(frames, meta["trans_input"], (S, S), out=out)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        preprocess.warp_normalize(frames, meta["trans_input"], (S, S), out=out)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[2]
    peaks, which = _peaks()
    nbytes = frames.numel() + out.numel() * 4
    gbs = nbytes / (ms / 1e3) / 1e9
    return {"bound": "hbm", "kernel": "preprocess_kernel", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
            "frac": gbs / peaks["hbm_gbs"], "traffic": None, "frames_per_launch": B, "ms": ms,
            "frames_per_s": B / (ms / 1e3), "peak_source": which}


def clip_pipeline(eng, dev, frames, sd=None):
    """BASELINE configs[3]-style pipeline on the bench batch: lock-step clips, device prior rendering +
    network + decode, host PnP (cv2) per clip between frames.  Synthetic detections (exact projections
    + 0.5 px noise) are planted before every step so that the host PnP leg runs for every clip (random-init
    heads detect almost nothing).  Host LM/PnP time is reported separately, as north_star asks."""
    import numpy as np
    from sgtapose_b200 import detector, synth
    det = detector.LockstepDetector(eng, workers=min(16, os.cpu_count() or 1))
    B = eng.B
    rng = np.random.default_rng(317)
    base = rng.uniform([-0.35, -0.2, 1.2], [0.35, 0.2, 1.8], size=(B, 7, 3))
    # RAW camera frames (uint8 640x360, what the reference's run() receives): uploaded as uint8 and
    # pre-processed on the device (warpAffine + normalise, sgta_detector.py:368-399) inside every step
    imgs = [torch.from_numpy(rng.integers(0, 256, (B, det.raw_h, det.raw_w, 3), dtype=np.uint8)).pin_memory()
            for f in range(2)]

    def x3d(f):
        return base + 0.004 * f

    def plant(f):
        p = np.einsum("ij,bkj->bki", det.K, x3d(f))
        det.detected_kps = p[:, :, :2] / p[:, :, 2:] + rng.normal(0, 0.5, size=(B, 7, 2))
    det.step(imgs[0])
    plant(0)
    det.step(imgs[1], x3d(0), x3d(1))                       # warm-up of the PnP path
    det.timing = {"host_pnp": 0.0, "host_post": 0.0, "steps": 0}
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for f in range(2, 2 + frames):
        plant(f - 1)
        det.step(imgs[f & 1], x3d(f - 1), x3d(f))
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    res = {"frames_per_s": B * frames / dt, "frames": frames, "clips": B, "ms_per_step": dt / frames * 1e3,
           "host_pnp_render_ms_per_step": det.timing["host_pnp"] / frames * 1e3,
           "host_post_ms_per_step": det.timing["host_post"] / frames * 1e3,
           "input": "raw uint8 %dx%d frames, device pre-processing" % (det.raw_w, det.raw_h),
           "h2d_bytes_per_step": imgs[0].numel() + 2 * 2 * B * 7 * 2 * 8, "d2h_bytes_per_step": B * 7 * 3 * 4}
    if sd is not None and B % 2 == 0:
        # the same B clips as two groups of B/2, and 2B clips as two groups of B (the bench batch per group)
        res["skewed_2_groups_half_batch"] = clip_groups_pipeline(sd, eng, B // 2, frames, rng)
        res["skewed_2_groups_full_batch"] = clip_groups_pipeline(sd, eng, B, frames, rng)
    return res


def clip_groups_pipeline(sd, eng, per_group, frames, rng):
    """Two lock-step groups of `per_group` clips (one engine each) run skewed
    (sgtapose_b200/detector.py::ClipGroups): the host PnP of one group runs under the device work of the other."""
    import numpy as np
    from sgtapose_b200 import detector, engine
    B = 2 * per_group
    workers = min(16, os.cpu_count() or 1)
    engs = [eng if per_group == eng.B and i == 0 else
            engine.InferenceEngine(sd, eng.opt, batch=per_group, size=eng.S, mode=eng.mode, device=eng.dev, fuse_sigmoid=True)
            for i in range(2)]
    groups = detector.ClipGroups([detector.LockstepDetector(e, workers=workers) for e in engs])
    K = groups.dets[0].K
    raw_h, raw_w = groups.dets[0].raw_h, groups.dets[0].raw_w
    base = rng.uniform([-0.35, -0.2, 1.2], [0.35, 0.2, 1.8], size=(B, 7, 3))
    imgs = [torch.from_numpy(rng.integers(0, 256, (B, raw_h, raw_w, 3), dtype=np.uint8)).pin_memory() for f in range(2)]

    def x3d(f):
        return base + 0.004 * f

    def plant(g, f, d):
        if d.frame == 0:
            return
        p = np.einsum("ij,bkj->bki", K, x3d(f - 1)[groups.offsets[g]:groups.offsets[g + 1]])
        d.detected_kps = p[:, :, :2] / p[:, :, 2:] + rng.normal(0, 0.5, size=(d.B, 7, 2))

    groups.run(3, lambda f: imgs[f & 1], x3d, before_begin=plant)       # warm-up (graphs, PnP path)
    for d in groups.dets:
        d.timing = {"host_pnp": 0.0, "host_post": 0.0, "steps": 0}
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    groups.run(frames, lambda f: imgs[f & 1], x3d, before_begin=plant)   # frame 0 of this run re-uses the kept state
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    host = sum(d.timing["host_pnp"] + d.timing["host_post"] for d in groups.dets)
    return {"frames_per_s": B * frames / dt, "ms_per_step": dt / frames * 1e3, "groups": 2, "clips_per_group": per_group,
            "clips": B, "host_ms_per_step_all_groups": host / frames * 1e3}


def time_dcn_kernels(fn):
    """Sum of the DCN kernel durations of one step, each launch bracketed by CUDA events."""
    from sgtapose_b200 import _lib
    events = []
    orig = _lib.call

    def timed(name, *a):
        if name in ("sgta_planes_dcn", "sgta_dcn_forward", "sgta_dcn_forward_nhwc"):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            orig(name, *a)
            e.record()
            events.append((s, e))
        else:
            orig(name, *a)

    _lib.call = timed
    try:
        tot = []
        for _ in range(3):
            events.clear()
            fn()
            torch.cuda.synchronize()
            tot.append(sum(s.elapsed_time(e) for s, e in events))
    finally:
        _lib.call = orig
    return sorted(tot)[1]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--skip-dead-levels", type=int, default=0)
    ap.add_argument("--engine", default="graph", choices=["graph", "eager"])
    ap.add_argument("--mode", default="fp32", choices=["fp32", "bf16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-sync", action="store_true", help="time e2e through the synchronous infer() call")
    ap.add_argument("--no-extras", action="store_true", help="skip the decode-roofline and clip-pipeline legs")
    ap.add_argument("--clip-frames", type=int, default=8)
    ap.add_argument("--dbg", type=int, default=0, help="sgta_debug_flags value (kernel experiments; 0 for any reported number)")
    ap.add_argument("--profile-pass", action="store_true",
                    help="run one un-graphed step inside cudaProfilerStart/Stop and exit (for ncu)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
