"""Load the UNMODIFIED host-side geometry / post-processing / metric code of the reference -- TEST INFRASTRUCTURE
(build container only; used by oracle/make_golden_r2.py and `reference`-marked CPU tests).

  geometric_vision.py   loaded whole by path.  Its two absent, un-pinned third-party imports are stubbed:
        `transforms3d` -> empty module (not on this path);
        `pyrr.Quaternion` -> `PyrrQuaternionStub` below, a restatement of the THREE pyrr members the path uses
        (`from_axis_rotation`, `normalize`, `matrix33`, `geometric_vision.py:15-25`, `:291`, `:189`) from pyrr's
        published formulas (pyrr 0.10.3 `quaternion.create_from_axis_rotation`, `quaternion.normalize`,
        `matrix33.create_from_quaternion`).  THE STUB IS OURS, NOT THE REFERENCE'S: what these goldens pin is the
        reference's control flow around it (cv2 EPnP -> iterative refinement, the `good` filter, the homogeneous
        transform and projection of `is_pnp`), and tests cross-check the stub's rotation against `cv2.Rodrigues`.
  lib/utils/post_process.py::dream_generic_post_process and lib/utils/image.py   loaded by path (relative imports
        resolved through a stub package; `ddd_utils` needs numba-free numpy only).
  lib/sgta_detector.py::{post_process, merge_outputs, _get_final_kps}   cut out with `ast` and compiled unchanged
        (the module itself pulls the whole detector stack in at import time), like oracle/make_golden_preprocess.py.
  analysis.py::{keypoint_metrics, pnp_metrics}   cut out with `ast` (module-level imports need matplotlib / ruamel).
"""
import ast
import importlib.util
import os
import sys
import types
from copy import deepcopy

import numpy as np

REF = os.environ.get("SGTA_REFERENCE_ROOT", "/root/reference")
PKG = os.path.join(REF, "sgtapose")


class PyrrQuaternionStub(np.ndarray):
    """xyzw quaternion with the pyrr members geometric_vision.py touches."""

    def __new__(cls, value):
        return np.asarray(value, dtype=np.float64).view(cls)

    @classmethod
    def from_axis_rotation(cls, axis, theta):
        axis = np.asarray(axis, dtype=np.float64)
        half = theta * 0.5
        s = np.sin(half)
        q = np.array([s * axis[0], s * axis[1], s * axis[2], np.cos(half)])
        return cls(q / np.sqrt(np.sum(q ** 2)))                 # create_from_axis_rotation normalises

    def normalize(self):
        self[:] = np.asarray(self) / np.sqrt(np.sum(np.asarray(self) ** 2))

    @property
    def matrix33(self):
        qx, qy, qz, qw = [float(v) for v in self]
        sqw, sqx, sqy, sqz = qw ** 2, qx ** 2, qy ** 2, qz ** 2
        invs = 1.0 / (sqx + sqy + sqz + sqw)
        qxy, qzw, qxz, qyw, qyz, qxw = qx * qy, qz * qw, qx * qz, qy * qw, qy * qz, qx * qw
        return np.array([[(sqx - sqy - sqz + sqw) * invs, 2.0 * (qxy - qzw) * invs, 2.0 * (qxz + qyw) * invs],
                         [2.0 * (qxy + qzw) * invs, (-sqx + sqy - sqz + sqw) * invs, 2.0 * (qyz - qxw) * invs],
                         [2.0 * (qxz - qyw) * invs, 2.0 * (qyz + qxw) * invs, (-sqx - sqy + sqz + sqw) * invs]])


def available():
    return os.path.isfile(os.path.join(PKG, "geometric_vision.py"))


def load_geometric_vision():
    pyrr = types.ModuleType("pyrr")
    pyrr.Quaternion = PyrrQuaternionStub
    saved = {k: sys.modules.get(k) for k in ("pyrr", "transforms3d")}
    sys.modules["pyrr"] = pyrr
    sys.modules["transforms3d"] = types.ModuleType("transforms3d")
    try:
        spec = importlib.util.spec_from_file_location("ref_geometric_vision", os.path.join(PKG, "geometric_vision.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod


def load_post_process():
    """-> (dream_generic_post_process, image module) from lib/utils, by path."""
    utils_dir = os.path.join(PKG, "lib", "utils")
    pkg = types.ModuleType("ref_lib_utils")
    pkg.__path__ = [utils_dir]
    sys.modules["ref_lib_utils"] = pkg
    mods = {}
    for name in ("image", "ddd_utils", "post_process"):
        spec = importlib.util.spec_from_file_location("ref_lib_utils." + name, os.path.join(utils_dir, name + ".py"))
        m = importlib.util.module_from_spec(spec)
        sys.modules["ref_lib_utils." + name] = m
        spec.loader.exec_module(m)
        mods[name] = m
    return mods["post_process"].dream_generic_post_process, mods["image"]


def _cut(path, names, cls=None, extra_ns=None):
    tree = ast.parse(open(path).read())
    body = tree.body
    if cls is not None:
        body = next(n for n in body if isinstance(n, ast.ClassDef) and n.name == cls).body
    keep = [n for n in body if isinstance(n, ast.FunctionDef) and n.name in names]
    assert len(keep) == len(names), (names, [n.name for n in keep])
    if cls is not None:
        keep = [ast.ClassDef(name=cls + "Cut", bases=[], keywords=[], body=keep, decorator_list=[])]
    mod = ast.Module(body=keep, type_ignores=[])
    ast.fix_missing_locations(mod)
    ns = dict(extra_ns or {})
    exec(compile(mod, path, "exec"), ns)
    return ns


def load_detector_post():
    """-> class with the reference's post_process / merge_outputs / _get_final_kps, needing .opt and .is_ct."""
    dgp, image_mod = load_post_process()
    ns = _cut(os.path.join(PKG, "lib", "sgta_detector.py"), ("post_process", "merge_outputs", "_get_final_kps"),
              cls="SGTADetector", extra_ns={"np": np, "deepcopy": deepcopy, "dream_generic_post_process": dgp})
    return ns["SGTADetectorCut"], image_mod


def load_metrics():
    """-> (keypoint_metrics, pnp_metrics) of analysis.py:1640-1793."""
    ns = _cut(os.path.join(PKG, "analysis.py"), ("keypoint_metrics", "pnp_metrics"), extra_ns={"np": np})
    return ns["keypoint_metrics"], ns["pnp_metrics"]
