"""Host-side pose refinement: the reference's `rf_tools/LM.py` entry points over
`sgta_lm_refine` (csrc/lm_refine.cu, host code, no device work -- north_star keeps LM/PnP on the host).

`register_GN_C`, `get_weights_without` and `get_weights` keep the reference's names, arguments and return
values (LM.py:256-266, :272-275, :277-307), so `analysis.py:202` runs unchanged on top of this module.
A maintainer can equally point LM.py:10 at `sgtapose_b200/libsgta_b200.so`: it exports `LM` with the
argument list of the reference's binary-only `libtestso_final.so`.
"""
import ctypes

import numpy as np

from . import _lib

_D = ctypes.POINTER(ctypes.c_double)


def _arr(a, n):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(-1))
    if a.size != n:
        raise _lib.SgtaError("register_GN_C: expected %d values, got %d" % (n, a.size))
    return a


def register_GN_C(x2d_input_p, x3d_input_p, quat_init, T_init, weights, camera, num_points):
    """quat_init (1,4) wxyz, T_init (1,3), x2d (num_points,2), x3d (num_points,3), weights (num_points+1,2),
    camera (3,3)  ->  (quat [4] wxyz, T [3]) as float64 arrays (NaN if the iteration broke down)."""
    num_points = int(num_points)
    v0 = _arr(np.hstack([np.asarray(quat_init, float).reshape(-1)[:4], np.asarray(T_init, float).reshape(-1)[:3]]), 7)
    x2d, x3d = _arr(x2d_input_p, 2 * num_points), _arr(x3d_input_p, 3 * num_points)
    w, cam = _arr(weights, 2 * num_points + 2), _arr(camera, 9)
    ans = np.zeros(7)
    _lib.call("sgta_lm_refine", v0.ctypes.data_as(_D), x2d.ctypes.data_as(_D), x3d.ctypes.data_as(_D),
              w.ctypes.data_as(_D), cam.ctypes.data_as(_D), ans.ctypes.data_as(_D), num_points)
    return ans[:4], ans[4:]


def get_weights_without(num_pt):
    """LM.py:272-275: unit weights, 1e8 for the unit-quaternion constraint row."""
    weights = np.ones((num_pt + 1, 2), dtype=float)
    weights[-1:] = 1e8
    return weights.tolist()


def get_weights(num_pt, distance):
    """LM.py:277-307 (the active choice, exp(-5 d)): distance [num_pt,2] squared reprojection distances."""
    weights = np.ones((num_pt + 1, 2), dtype=float)
    weights[:num_pt] = np.exp(-5 * np.asarray(distance, float)[:num_pt])
    weights[-1:] = 1e8
    return weights.tolist()


def get_weights_real(x2d_input, x3d_input, transform, camera):
    """LM.py:309-332: weights from the squared reprojection distance of every coordinate under `transform` (3x4 or 4x4
    pose): 0 beyond 100 px^2, 1 below 1 px^2, 1000^(1 - d/10) / 1000 between; points flagged missing (x < -1000) keep
    weight 0.  -> (weights [num_points+1, 2] ndarray, num_points)."""
    x2d_input, x3d_input = np.asarray(x2d_input, float), np.asarray(x3d_input, float)
    num_points = x2d_input.shape[0]
    weights = np.zeros((num_points + 1, 2))
    P = np.asarray(camera, float) @ np.asarray(transform, float)[0:3]
    for i in range(num_points):
        if x2d_input[i, 0] < -1000:
            continue
        rep = P @ np.append(x3d_input[i], 1.0)
        dis = (rep[:2] / rep[2] - x2d_input[i]) ** 2
        for j in range(2):
            if dis[j] > 100:
                weights[i, j] = 0
            elif dis[j] < 1:
                weights[i, j] = 1
            else:
                weights[i, j] = np.power(1000, (1 - (dis[j] / 10))) / 1000
    weights[-1] = [1e8, 1e8]
    return weights, num_points
