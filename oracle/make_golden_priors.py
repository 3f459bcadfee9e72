"""Generate tests/golden/priors.npz from the UNMODIFIED reference sgtapose/utilities.py.

Run in the build container only:  python -m oracle.make_golden_priors
utilities.py:9 imports ruamel.yaml (absent here, unused on this path) -> stubbed.
Cases cover: keypoints inside the frame, outside the raw image (-> (0,0), nothing drawn),
closer than radius+1 to a map border (nothing drawn), overlapping blobs (max-blend), points
that land exactly on integer map coordinates, and kp_projs_raw = None (all-zero maps).
"""
import importlib.util
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden", "priors.npz")
REF = os.environ.get("SGTA_REFERENCE_ROOT", "/root/reference")

RAW_W, RAW_H = 640, 360


def load_utilities():
    if "ruamel" not in sys.modules:
        ru, ry = types.ModuleType("ruamel"), types.ModuleType("ruamel.yaml")
        ry.YAML = object
        ru.yaml = ry
        sys.modules["ruamel"], sys.modules["ruamel.yaml"] = ru, ry
    spec = importlib.util.spec_from_file_location("ref_utilities", os.path.join(REF, "sgtapose", "utilities.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def cases(S):
    """Raw-image keypoint sets [n_case, 7, 2] (float64) for an S x S network input."""
    rng = np.random.default_rng(1000 + S)
    out = []
    out.append(rng.uniform([40, 40], [RAW_W - 40, RAW_H - 40], size=(7, 2)))            # all inside
    a = rng.uniform([40, 40], [RAW_W - 40, RAW_H - 40], size=(7, 2))
    a[1] = [-3.0, 100.0]; a[3] = [700.0, 20.0]; a[5] = [320.0, 360.0]                    # outside raw
    out.append(a)
    b = rng.uniform([40, 40], [RAW_W - 40, RAW_H - 40], size=(7, 2))
    b[0] = [1.0, 180.0]; b[2] = [638.9, 180.0]; b[4] = [320.0, 0.5]; b[6] = [320.0, 359.5]  # near borders
    out.append(b)
    c = np.tile(np.array([[300.0, 200.0]]), (7, 1)) + rng.uniform(-6, 6, size=(7, 2))  # overlapping
    out.append(c)
    d = np.array([[5.0 * k / 3.0 * 10, 100.0 + 5.0 * k] for k in range(1, 8)])         # integer grid at S=384
    out.append(d)
    out.append(rng.uniform([0, 0], [RAW_W, RAW_H], size=(7, 2)))                        # anywhere
    return np.stack(out)


def main():
    U = load_utilities()
    blob = {}
    for S in (128, 384):
        q = S // 4
        c = np.array([RAW_W / 2.0, RAW_H / 2.0], dtype=np.float32)        # sgta_detector.py:354-357 (fix_res)
        s = max(RAW_H, RAW_W) * 1.0
        t_in = U.get_affine_transform(c, s, 0, [S, S])
        t_out = U.get_affine_transform(c, s, 0, [q, q])
        kps = cases(S)
        hms, clss, cin, cout = [], [], [], []
        for kp in kps:
            hms.append(U.get_prev_hm_wo_noise(kp, t_in, S, S, RAW_W, RAW_H))
            clss.append(U.get_prev_hm_wo_noise_cls(kp, kp, t_out, q, q, RAW_W, RAW_H))
            cin.append(U.affine_transform_and_clip(kp, t_in, S, S, RAW_W, RAW_H))
            cout.append(U.affine_transform_and_clip(kp, t_out, q, q, RAW_W, RAW_H))
        blob["S%d_kps" % S] = kps
        blob["S%d_trans_input" % S] = t_in
        blob["S%d_trans_output" % S] = t_out
        blob["S%d_centres_in" % S] = np.stack(cin)
        blob["S%d_centres_out" % S] = np.stack(cout)
        # the maps are sparse: store the non-zero entries only (flat index, value)
        hm, cl = np.stack(hms), np.stack(clss)
        for name, arr in (("hm", hm), ("cls", cl)):
            nz = np.flatnonzero(arr)
            blob["S%d_%s_shape" % (S, name)] = np.array(arr.shape)
            blob["S%d_%s_idx" % (S, name)] = nz.astype(np.int64)
            blob["S%d_%s_val" % (S, name)] = arr.reshape(-1)[nz]
        blob["S%d_none_hm_sum" % S] = np.array(U.get_prev_hm_wo_noise(None, t_in, S, S, RAW_W, RAW_H).sum())
    blob["gaussian"] = U.gaussian2D((9, 9), sigma=2, res=[0, 0])
    np.savez_compressed(OUT, **blob)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
