"""CPU restatement of the reference's pose refinement -- TEST INFRASTRUCTURE (only tests/ and the golden
generator import this).

Reference: sgtapose/rf_tools/LM.py.  `register_GN_C` (:256-266) calls `LM` of the binary-only
`libtestso_final.so`; the same file's Python twin states what it computes:
  fun   :128-156   residual vector F (2n+1): w^2 (x2d - u)^2, w^2 (y2d - v)^2 per point, then the
                   unit-quaternion term 1e8 s^2 + 1e8 s^2, s = |q|^2 - 1
  dfun  :161-217   its Jacobian (generated expressions); the d F_v / d qz entry (:208) is the d F_v / d qy
                   expression (:207) repeated
  GN    :220-232   value -= inv(J^T J + 1e-4 I) J^T F until sum|delta| <= 1e-4 or 200 iterations
Probed against the binary (oracle/make_golden_lm.py): the .so runs that iteration in float64 (the twin
casts J^T J to float32, the .so does not) with the Jacobian exactly as written.  Restated here with the
Jacobian derived from the quaternion product; pinned to the twin's `fun` / `dfun` (tests/test_lm.py, build
container) and to the .so's outputs (tests/golden/lm.npz).
"""
import numpy as np


def _points(v, x3d):
    qw, qx, qy, qz, tx, ty, tz = v
    x, y, z = x3d[:, 0], x3d[:, 1], x3d[:, 2]
    a = qw * x + qy * z - qz * y
    b = qw * y - qx * z + qz * x
    c = qw * z + qx * y - qy * x
    d = -qx * x - qy * y - qz * z
    Px = qw * a - qx * d + qy * c - qz * b + tx
    Py = qw * b - qx * c - qy * d + qz * a + ty
    Pz = qw * c + qx * b - qy * a - qz * d + tz
    return a, b, c, d, Px, Py, Pz


def fun(v, x2d, x3d, weights, camera):
    x2d, x3d, w = np.asarray(x2d, float), np.asarray(x3d, float), np.asarray(weights, float)
    fx, cx, fy, cy = camera[0][0], camera[0][2], camera[1][1], camera[1][2]
    n = len(x2d)
    _, _, _, _, Px, Py, Pz = _points(v, x3d)
    F = np.empty(2 * n + 1)
    F[0:2 * n:2] = w[:n, 0] ** 2 * (x2d[:, 0] - (cx * Pz + fx * Px) / Pz) ** 2
    F[1:2 * n:2] = w[:n, 1] ** 2 * (x2d[:, 1] - (cy * Pz + fy * Py) / Pz) ** 2
    s = v[0] ** 2 + v[1] ** 2 + v[2] ** 2 + v[3] ** 2 - 1
    F[2 * n] = 1e8 * s ** 2 + 1e8 * s ** 2
    return F


def dfun(v, x2d, x3d, weights, camera):
    x2d, x3d, w = np.asarray(x2d, float), np.asarray(x3d, float), np.asarray(weights, float)
    fx, cx, fy, cy = camera[0][0], camera[0][2], camera[1][1], camera[1][2]
    n = len(x2d)
    a, b, c, d, Px, Py, Pz = _points(v, x3d)
    ru = x2d[:, 0] - (cx * Pz + fx * Px) / Pz
    rv = x2d[:, 1] - (cy * Pz + fy * Py) / Pz
    dPx = np.stack([2 * a, -2 * d, 2 * c, -2 * b], 1)
    dPy = np.stack([2 * b, -2 * c, -2 * d, 2 * a], 1)
    dPz = np.stack([2 * c, 2 * b, -2 * a, -2 * d], 1)
    dPy_ref, dPz_ref = dPy.copy(), dPz.copy()
    dPy_ref[:, 3], dPz_ref[:, 3] = dPy[:, 2], dPz[:, 2]          # LM.py:208 repeats :207
    J = np.zeros((2 * n + 1, 7))
    wx2, wy2 = w[:n, 0] ** 2, w[:n, 1] ** 2
    du = fx * (dPx * Pz[:, None] - Px[:, None] * dPz) / Pz[:, None] ** 2
    dv = fy * (dPy_ref * Pz[:, None] - Py[:, None] * dPz_ref) / Pz[:, None] ** 2
    J[0:2 * n:2, :4] = (-2 * wx2 * ru)[:, None] * du
    J[1:2 * n:2, :4] = (-2 * wy2 * rv)[:, None] * dv
    J[0:2 * n:2, 4] = -2 * wx2 * ru * fx / Pz
    J[0:2 * n:2, 6] = -2 * wx2 * ru * (-fx * Px / Pz ** 2)
    J[1:2 * n:2, 5] = -2 * wy2 * rv * fy / Pz
    J[1:2 * n:2, 6] = -2 * wy2 * rv * (-fy * Py / Pz ** 2)
    s = v[0] ** 2 + v[1] ** 2 + v[2] ** 2 + v[3] ** 2 - 1
    J[2 * n, :4] = 8e8 * s * np.asarray(v[:4])                    # 4 q wx s + 4 q wy s with wx = wy = 1e8 (:222-229)
    return J


def gn(value, x2d, x3d, weights, camera, max_iter=200):
    """-> (value [7], iterations)."""
    value = np.asarray(value, float).copy()
    delta = np.ones(7) * 100
    i = 0
    while np.sum(np.abs(delta)) > 1e-4 and i < max_iter:
        J = dfun(value, x2d, x3d, weights, camera)
        F = fun(value, x2d, x3d, weights, camera)
        value1 = value - np.linalg.inv(J.T @ J + 1e-4 * np.identity(7)) @ J.T @ F
        delta = value1 - value
        value = value1
        i += 1
    return value, i


def get_weights_without(num_pt):
    """LM.py:272-275."""
    w = np.ones((num_pt + 1, 2), dtype=float)
    w[-1:] = 1e8
    return w.tolist()


def rotation_from_quaternion(q):
    """LM.py:92-108 compute_rotation_matric_from_quaternion (wxyz, normalised first)."""
    qw, qx, qy, qz = np.asarray(q, float) / np.linalg.norm(q)
    return np.array([[1 - 2 * qy * qy - 2 * qz * qz, 2 * qx * qy - 2 * qz * qw, 2 * qx * qz + 2 * qy * qw],
                     [2 * qx * qy + 2 * qz * qw, 1 - 2 * qx * qx - 2 * qz * qz, 2 * qy * qz - 2 * qx * qw],
                     [2 * qx * qz - 2 * qy * qw, 2 * qy * qz + 2 * qx * qw, 1 - 2 * qx * qx - 2 * qy * qy]])
