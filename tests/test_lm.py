"""Host-side pose refinement (sgtapose_b200/lm.py -> sgta_lm_refine / LM) vs the outputs of the reference's
own libtestso_final.so (tests/golden/lm.npz), the oracle restatement, and -- in the build container -- the
reference's Python twin.  CPU tests: the entry is host code."""
import ast
import ctypes
import os

import numpy as np
import pytest

from oracle import lm as olm
from oracle import ref_import

WELL = 40        # problems the float64 iteration finishes in <= WELL steps are compared tightly; the iteration
                 # stops on step size and wanders on the others, where 1e-16 arithmetic differences grow to 1e-3


def _problems(golden):
    g = golden("lm.npz")
    i = 0
    while "n_%d" % i in g:
        yield i, int(g["n_%d" % i]), g["v0_%d" % i], g["x2d_%d" % i], g["x3d_%d" % i], g["w_%d" % i], g["ans_%d" % i], \
            int(g["its_%d" % i]), g["camera"]
        i += 1


def _pose_error(a, b):
    """(translation error [m], rotation error [deg]) between two (wxyz, t) 7-vectors."""
    Ra, Rb = olm.rotation_from_quaternion(a[:4]), olm.rotation_from_quaternion(b[:4])
    c = np.clip((np.trace(Ra.T @ Rb) - 1) / 2, -1, 1)
    return float(np.linalg.norm(a[4:] - b[4:])), float(np.degrees(np.arccos(c)))


def test_oracle_gn_matches_reference_binary(golden):
    n_well = 0
    for i, n, v0, x2d, x3d, w, ans, its, K in _problems(golden):
        mine, my_its = olm.gn(v0, x2d, x3d, w, K)
        assert my_its == its
        if not np.isfinite(ans).all():                            # the iteration blew up in the binary too
            assert np.isnan(mine).any(), i
        elif its <= WELL:
            n_well += 1
            assert np.abs(mine - ans).max() < 1e-5, (i, np.abs(mine - ans).max())
    assert n_well >= 20


def test_lm_refine_matches_reference_binary(golden):
    from sgtapose_b200 import lm
    for i, n, v0, x2d, x3d, w, ans, its, K in _problems(golden):
        q, t = lm.register_GN_C(x2d.tolist(), x3d.tolist(), v0[:4].reshape(1, 4), v0[4:].reshape(1, 3), w.tolist(), K, n)
        got = np.hstack([q, t])
        if not np.isfinite(ans).all():
            # blow-up: the caller tests isnan(quat) / isnan(T) and keeps the PnP pose (analysis.py:206-210)
            assert np.isnan(got).any(), i
        elif its <= WELL:
            assert np.abs(got - ans).max() < 1e-5, (i, np.abs(got - ans).max())
            dt, dr = _pose_error(got, ans)
            assert dt < 1e-3 and dr < 0.1, (i, dt, dr)            # north_star: 1 mm / 0.1 deg
            mine, _ = olm.gn(v0, x2d, x3d, w, K)
            assert np.abs(got - mine).max() < 1e-5, i


def test_reference_symbol_LM_is_exported(golden):
    """LM.py:10 / :264 bind `so.LM(value_init, x2d, x3d, weights, camera, ans, num_points)`."""
    from sgtapose_b200 import _lib
    so = ctypes.CDLL(_lib.LIB_PATH)
    i, n, v0, x2d, x3d, w, ans, its, K = next(p for p in _problems(golden) if p[7] <= WELL and np.isfinite(p[6]).all())
    D = ctypes.c_double
    out = (D * 7)(*([0.0] * 7))
    so.LM((D * 7)(*v0), (D * (2 * n))(*x2d.reshape(-1)), (D * (3 * n))(*x3d.reshape(-1)),
          (D * (2 * n + 2))(*w.reshape(-1)), (D * 9)(*K.reshape(-1)), out, n)
    assert np.abs(np.array(list(out)) - ans).max() < 1e-5
    # bad sizes: error code from the sgta_ entry
    from sgtapose_b200 import lm
    with pytest.raises(_lib.SgtaError):
        lm.register_GN_C(np.zeros((0, 2)), np.zeros((0, 3)), np.ones((1, 4)), np.ones((1, 3)), np.ones((1, 2)), K, 0)


def test_lm_many_points_and_guard_only_in_sgta_entry(golden):
    """No cap on the number of correspondences (the reference has none): a 70-point problem built by repeating a
    golden problem's points converges to the same pose.  The behind-the-camera guard lives in `sgta_lm_refine` only:
    started from a mirror pose the drop-in `LM` symbol returns its finite iterate, like the reference binary."""
    from sgtapose_b200 import _lib, lm
    i, n, v0, x2d, x3d, w, ans, its, K = next(p for p in _problems(golden) if p[7] <= WELL and np.isfinite(p[6]).all())
    rep = 12
    w_rep = np.vstack([np.tile(w[:n], (rep, 1)) / rep, w[n:]])
    q, T = lm.register_GN_C(np.tile(x2d, (rep, 1)), np.tile(x3d, (rep, 1)), v0[None, :4], v0[None, 4:], w_rep, K, n * rep)
    assert n * rep >= 70 and np.isfinite(q).all() and np.abs(np.concatenate([q, T]) - ans).max() < 1e-3
    so = ctypes.CDLL(_lib.LIB_PATH)
    D = ctypes.c_double
    mirror = v0.copy()
    mirror[6] = -abs(mirror[6])                                   # start behind the camera, tiny step budget
    out = (D * 7)(*([0.0] * 7))
    so.LM((D * 7)(*mirror), (D * (2 * n))(*x2d.reshape(-1)), (D * (3 * n))(*x3d.reshape(-1)),
          (D * (2 * n + 2))(*w.reshape(-1)), (D * 9)(*K.reshape(-1)), out, n)
    raw = np.array(list(out))
    q2, T2 = lm.register_GN_C(x2d, x3d, mirror[None, :4], mirror[None, 4:], w, K, n)
    guarded = np.concatenate([q2, T2])
    # wherever the guard fired the un-guarded entry still reports its finite iterate; elsewhere they are identical
    if np.isnan(guarded).any() and np.isfinite(raw).all():
        assert True
    else:
        assert np.array_equal(np.nan_to_num(guarded, nan=-1.0), np.nan_to_num(raw, nan=-1.0))


def test_weights_helpers():
    from sgtapose_b200 import lm
    assert lm.get_weights_without(3) == olm.get_weights_without(3) == [[1.0, 1.0]] * 3 + [[1e8, 1e8]]
    d = np.array([[0.0, 0.2], [1.0, 0.1]])
    w = np.array(lm.get_weights(2, d))
    assert np.allclose(w[:2], np.exp(-5 * d)) and w[2].tolist() == [1e8, 1e8]


@pytest.mark.reference
def test_oracle_residuals_and_jacobian_match_reference_twin():
    """oracle/lm.py fun / dfun vs the UNMODIFIED `fun` / `dfun` of rf_tools/LM.py (cut out with ast: the module
    itself loads a binary from a hard-coded absolute path at import, LM.py:10)."""
    path = os.path.join(ref_import.REF_PKG, "rf_tools", "LM.py")
    tree = ast.parse(open(path).read())
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("fun", "dfun", "get_weights_without")]
    mod = ast.Module(body=keep, type_ignores=[])
    ast.fix_missing_locations(mod)
    ns = {"np": np}
    exec(compile(mod, path, "exec"), ns)
    rng = np.random.default_rng(4)
    K = np.array([[502.30, 0.0, 319.75], [0.0, 502.30, 179.75], [0.0, 0.0, 1.0]])
    for n in (4, 7):
        v = np.hstack([rng.normal(size=4), rng.uniform(-0.2, 0.2, 2), rng.uniform(1, 2, 1)])
        x3d, x2d = rng.uniform(-0.4, 0.4, (n, 3)), rng.uniform(0, 600, (n, 2))
        w = np.array(ns["get_weights_without"](n))
        w[:n] = rng.uniform(0.3, 1.0, (n, 2))
        F_ref = np.array(ns["fun"](v, x2d.tolist(), x3d.tolist(), w.tolist(), K), float)
        J_ref = np.array(ns["dfun"](v, x2d.tolist(), x3d.tolist(), w.tolist(), K), float)
        np.testing.assert_allclose(olm.fun(v, x2d, x3d, w, K), F_ref, rtol=1e-10, atol=1e-9)
        np.testing.assert_allclose(olm.dfun(v, x2d, x3d, w, K), J_ref, rtol=1e-9, atol=1e-6)


@pytest.mark.reference
def test_get_weights_real_matches_reference():
    from sgtapose_b200 import lm
    path = os.path.join(ref_import.REF_PKG, "rf_tools", "LM.py")
    tree = ast.parse(open(path).read())
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("get_weights_real", "get_weights")]
    mod = ast.Module(body=keep, type_ignores=[])
    ast.fix_missing_locations(mod)
    ns = {"np": np}
    exec(compile(mod, path, "exec"), ns)
    rng = np.random.default_rng(9)
    K = np.array([[502.30, 0.0, 319.75], [0.0, 502.30, 179.75], [0.0, 0.0, 1.0]])
    T = np.eye(4)
    T[:3, :3] = olm.rotation_from_quaternion(rng.normal(size=4))
    T[:3, 3] = [0.1, -0.05, 1.5]
    X = rng.uniform(-0.3, 0.3, (7, 3))
    uv = (K @ (T[:3] @ np.c_[X, np.ones(7)].T)).T
    uv = uv[:, :2] / uv[:, 2:] + rng.choice([0.3, 2.0, 15.0], size=(7, 2)) * rng.choice([-1, 1], size=(7, 2))
    uv[3] = [-5000.0, -5000.0]                                   # flagged missing
    w_ref, n_ref = ns["get_weights_real"](uv, X, T, K)
    w, n = lm.get_weights_real(uv, X, T, K)
    assert n == n_ref == 7 and np.array_equal(w, w_ref)
    d = rng.uniform(0, 1, (7, 2))
    assert np.allclose(np.array(lm.get_weights(7, d)), np.array(ns["get_weights"](7, d)), rtol=1e-15, atol=0)
