"""Round-2 fixtures, generated from the UNMODIFIED reference in the build container:  python -m oracle.make_golden_r2

  model_S384_seed{0,317}.npz   whole-network outputs at the headline size (384x384, B=2), reference modules with the
                               torchvision DCN stand-in (oracle/ref_import.py), SURVEY.md 8c(3)
  pnp.npz                      geometric_vision.is_pnp / solve_pnp (:43-116, :283-310) on planted detections
  pose.npz                     synthetic heat maps -> reference dream_generic_decode -> post_process / merge_outputs /
                               _get_final_kps -> solve_pnp: raw-pixel keypoints and poses (xyz + xyzw), SURVEY.md H9
  metrics.npz                  analysis.keypoint_metrics / pnp_metrics (:1640-1793), geometric_vision.add_from_pose
pyrr is absent: its three members on this path are stubbed (oracle/ref_geom.py says exactly how).
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_geom as G             # noqa: E402
from oracle import ref_import as R           # noqa: E402
from sgtapose_b200 import synth              # noqa: E402
from tests import _cases as C                # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def model_384(ns, opt):
    # the reference's write-back `cur_f[batch_id, feat_id] = mlp(...)` (dla.py:1014) has duplicate
    # indices whenever two keypoints share a cell (always possible at levels 3-5), and torch's CPU index_put_ is a
    # RACE on duplicates when it runs multi-threaded (observed here at 384x384, seed 317, sample 1, level 5: the
    # pixel got 172 channels of token 2 and 340 of token 3).  Single-threaded it is sequential = "highest token
    # index wins", the rule the oracle and the product define; the goldens are generated that way.
    torch.set_num_threads(1)
    model = R.build_reference_model(ns, opt)
    for seed in (0, 317):
        sd = synth.synthetic_state_dict(model.state_dict(), seed=seed)
        model.load_state_dict(sd)
        ins = synth.synthetic_inputs(2, 384, seed=seed, frame=1)
        out = model(*ins)[0]
        # the SAME unmodified modules in float64: the exact result the float32 runs scatter around.  The synthetic
        # weights make the 16-deep DeformConv chain ill-conditioned at this size: the reference's own float32 run is
        # 5e-4 (seed 317) / 4e-3 (seed 0) away from it, and differs from itself by as much between thread counts.
        import copy
        out64 = copy.deepcopy(model).double()(*[t.double() for t in ins])[0]
        path = os.path.join(OUT, "model_S384_seed%d.npz" % seed)
        blobs = {k: out[k].numpy() for k in ("hm", "reg", "tracking")}
        blobs.update({k + "_f64": out64[k].numpy().astype(np.float32) for k in ("hm", "reg", "tracking")})
        np.savez_compressed(path, **blobs)
        rel = lambda a, b: float((a.double() - b.double()).abs().max() / b.double().abs().max())
        print(path, os.path.getsize(path), "ref32 vs ref64:", {k: rel(out[k], out64[k]) for k in blobs if not k.endswith("_f64")})


def pnp(gv):
    blobs = {}
    for i, (prev, kps, nxt) in enumerate(C.pnp_cases()):
        good = np.unique(np.where(kps > C.MISSING)[0])
        a, b = gv.is_pnp(prev[good], kps[good], nxt, kps, C.CAMERA_K)
        ok, t, q = gv.solve_pnp(prev[good], kps[good], C.CAMERA_K)
        blobs["prev_%d" % i], blobs["next_%d" % i] = np.asarray(a, np.float64), np.asarray(b, np.float64)
        blobs["ok_%d" % i] = np.array(bool(ok))
        if ok:
            blobs["t_%d" % i], blobs["q_%d" % i] = np.asarray(t, np.float64), np.asarray(q, np.float64)
            blobs["R_%d" % i] = np.asarray(q.matrix33, np.float64)
        print("pnp", i, bool(ok), None if not ok else np.round(t, 4))
    np.savez_compressed(os.path.join(OUT, "pnp.npz"), **blobs)


def pose(ns, opt, gv):
    hm, reg, trk, x3d = C.pose_heatmap_cases()
    Det, image_mod = G.load_detector_post()
    det = Det()
    # opts_parallel.py:212, :263, :333: out_thresh = max(track_thresh = 0.001, out_thresh = -1)
    det.opt = types.SimpleNamespace(out_thresh=0.001, num_classes=7, test_scales=[1])
    det.is_ct = True
    q = hm.shape[2]
    meta = {"c": np.array([C.RAW_W / 2.0, C.RAW_H / 2.0], dtype=np.float32), "s": max(C.RAW_H, C.RAW_W) * 1.0,
            "out_height": q, "out_width": q, "height": C.RAW_H, "width": C.RAW_W, "calib": None}
    kps_all, poses, oks = [], [], []
    for n in range(hm.shape[0]):
        o = {"hm": torch.from_numpy(hm[n:n + 1]), "reg": torch.from_numpy(reg[n:n + 1]),
             "tracking": torch.from_numpy(trk[n:n + 1])}
        dets = ns.decode.dream_generic_decode(o, K=7, opt=opt)
        dets = {k: v.detach().cpu().numpy() for k, v in dets.items()}            # sgta_detector.py:922-924
        res = det.merge_outputs([det.post_process(dets, meta, 1)])
        kps = det._get_final_kps(res)
        good = np.unique(np.where(kps > -999.0)[0])                              # analysis.py:802-806
        ok, t, quat = gv.solve_pnp(x3d[n][good], kps[good], C.CAMERA_K)
        kps_all.append(kps)
        oks.append(bool(ok))
        poses.append(np.concatenate([t, np.asarray(quat)]) if ok else np.full(7, -999.99))
        print("pose", n, len(good), bool(ok), np.round(poses[-1], 4))
    np.savez_compressed(os.path.join(OUT, "pose.npz"), kps_raw=np.stack(kps_all), pose_xyz_xyzw=np.stack(poses),
                        ok=np.array(oks))


def metrics(gv):
    det, gt, add, inframe = C.metrics_case()
    km, pm = G.load_metrics()
    blobs = {}
    for syn in (False, True):
        r = km(det, gt, np.zeros(len(gt)), (C.RAW_W, C.RAW_H), 12.0, syn)
        for k, v in r.items():
            blobs["kp_%d_%s" % (syn, k)] = np.float64(v)
    r = pm(add, inframe)
    for k, v in r.items():
        blobs["pnp_%s" % k] = np.float64(v)
    rng = np.random.default_rng(5)
    x3d = C.panda_scene(rng, 4)
    adds = []
    for i in range(4):
        kps = C.project(x3d[i]) + rng.normal(0, 1.0, size=(7, 2))
        ok, t, q = gv.solve_pnp(x3d[i], kps, C.CAMERA_K)
        adds.append(gv.add_from_pose(t, q, x3d[i], C.CAMERA_K))
        blobs["add_t_%d" % i], blobs["add_q_%d" % i] = np.asarray(t), np.asarray(q)
    blobs["add_values"] = np.array(adds)
    np.savez_compressed(os.path.join(OUT, "metrics.npz"), **blobs)
    print("metrics", {k: float(v) for k, v in blobs.items() if k.startswith("pnp_add")}, adds)


def main():
    which = set(sys.argv[1:]) or {"pnp", "pose", "metrics", "model"}
    torch.set_grad_enabled(False)
    torch.Tensor.cuda = lambda self, *a, **k: self
    ns = R.load_reference()
    opt = R.default_opt()
    gv = G.load_geometric_vision()
    if "pnp" in which:
        pnp(gv)
    if "pose" in which:
        pose(ns, opt, gv)
    if "metrics" in which:
        metrics(gv)
    if "model" in which:
        model_384(ns, opt)


if __name__ == "__main__":
    main()
