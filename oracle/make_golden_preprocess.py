"""Generate tests/golden/preprocess.npz from the UNMODIFIED reference pre-processing code.

Run in the build container only:  python -m oracle.make_golden_preprocess
`sgtapose/lib/sgta_detector.py` cannot be imported here (progress, ruamel.yaml, PIL and the whole
detector stack are pulled in at module level, :15-32), so the three methods on this path --
`SGTADetector._transform_scale` (:334-366), `pre_process` (:368-399) and `normalize_img` (:402-403) --
are cut out of the source file with `ast`, compiled UNCHANGED, and bound to a stub object that carries
only the attributes they read (`opt`, `mean`, `std`, `rest_focal_length`).  `get_affine_transform`
comes from the reference's own `lib/utils/image.py`, loaded by path.  cv2 is the build container's
(opencv-python 4.13.0; the reference leaves it un-pinned, requirements.txt:7).

Cases: the BASELINE raw frame (640x360 -> 384x384 and the repo default 480x480), a portrait frame,
a frame smaller than the network input (up-sampling, border rows), odd sizes.
"""
import ast
import importlib.util
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden", "preprocess.npz")
REF = os.environ.get("SGTA_REFERENCE_ROOT", "/root/reference")

CASES = [  # raw (h, w), network input (h, w), seed
    ((360, 640), (384, 384), 1),
    ((360, 640), (480, 480), 2),
    ((640, 360), (384, 384), 3),
    ((120, 200), (384, 384), 4),
    ((357, 501), (320, 448), 5),
]


def case_image(hw, seed):
    """Smooth structure + noise + saturated patches, uint8 HWC (what cv2.imread returns)."""
    rng = np.random.default_rng(seed)
    h, w = hw
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    img = np.stack([127 + 120 * np.sin(xx / 17.0 + c) * np.cos(yy / 23.0 - c) for c in range(3)], axis=2)
    img += rng.normal(0, 12, size=img.shape)
    img[: h // 7, : w // 5] = 255
    img[-(h // 9):, -(w // 6):] = 0
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def load_reference_methods():
    import cv2
    import torch
    spec = importlib.util.spec_from_file_location("ref_image", os.path.join(REF, "sgtapose", "lib", "utils", "image.py"))
    image_mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(image_mod)
    path = os.path.join(REF, "sgtapose", "lib", "sgta_detector.py")
    tree = ast.parse(open(path).read())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "SGTADetector")
    keep = [n for n in cls.body if isinstance(n, ast.FunctionDef)
            and n.name in ("_transform_scale", "pre_process", "normalize_img", "_get_default_calib")]
    assert len(keep) == 4
    mod = ast.Module(body=[ast.ClassDef(name="RefPre", bases=[], keywords=[], body=keep, decorator_list=[])],
                     type_ignores=[])
    ast.fix_missing_locations(mod)
    ns = {"np": np, "cv2": cv2, "torch": torch, "get_affine_transform": image_mod.get_affine_transform}
    exec(compile(mod, path, "exec"), ns)
    return ns["RefPre"], torch


MODE_CASES = [  # raw (h, w), opt overrides, seed: the two other testing modes of _transform_scale (meta only is pinned
    ((360, 640), {"fix_short": 384}, 6),        # on the device side: the warp kernel is mode-independent)
    ((640, 360), {"fix_short": 320}, 7),
    ((357, 501), {"fix_res": False}, 8),
]


def reference_pre_process(image, input_hw, **over):
    RefPre, torch = load_reference_methods()
    obj = RefPre()
    obj.opt = types.SimpleNamespace(fix_short=-1, fix_res=True, input_h=input_hw[0], input_w=input_hw[1],
                                    down_ratio=4, pad=31)
    for k, v in over.items():
        setattr(obj.opt, k, v)
    obj.mean = torch.tensor([0.5, 0.5, 0.5], dtype=torch.float32).reshape(1, 1, 3)     # sgta_detector.py:58-59
    obj.std = torch.tensor([0.5, 0.5, 0.5], dtype=torch.float32).reshape(1, 1, 3)
    obj.rest_focal_length = 502.30
    images, meta = obj.pre_process(image, 1, {})
    return images.numpy(), meta


def main():
    out = {}
    for i, (raw, inp, seed) in enumerate(CASES):
        img = case_image(raw, seed)
        images, meta = reference_pre_process(img, inp)
        out["images_%d" % i] = images
        out["trans_input_%d" % i] = np.asarray(meta["trans_input"], np.float64)
        out["trans_output_%d" % i] = np.asarray(meta["trans_output"], np.float64)
        out["c_%d" % i] = np.asarray(meta["c"], np.float32)
        out["s_%d" % i] = np.float64(meta["s"])
        print(i, raw, inp, images.shape, float(images.min()), float(images.max()))
    for j, (raw, over, seed) in enumerate(MODE_CASES):
        img = case_image(raw, seed)
        images, meta = reference_pre_process(img, (384, 384), **over)
        out["mode_images_%d" % j] = images[:, :, ::4, ::4].copy()      # every 4th pixel: keeps the fixture small
        out["mode_trans_input_%d" % j] = np.asarray(meta["trans_input"], np.float64)
        out["mode_trans_output_%d" % j] = np.asarray(meta["trans_output"], np.float64)
        out["mode_sizes_%d" % j] = np.array([meta["inp_height"], meta["inp_width"], meta["out_height"], meta["out_width"]])
        print("mode", j, raw, over, images.shape)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
