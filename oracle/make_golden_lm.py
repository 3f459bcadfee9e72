"""Generate tests/golden/lm.npz from the reference's OWN binary `rf_tools/libtestso_final.so` (the entry
`register_GN_C` calls, LM.py:256-266) on seeded Panda-like problems.  Build container only:
    python -m oracle.make_golden_lm
Each problem: 4..7 keypoints of a robot-sized point set in front of the DEFAULT camera (sgta_detector.py:83),
detections = exact projections + Gaussian pixel noise, start = true pose perturbed like a PnP estimate, weights
from `get_weights_without` (LM.py:272-275) or random values in (0.5, 1].  Stored per problem: the inputs, the
.so's answer and the iteration count of the float64 restatement (oracle/lm.py; 200 = did not converge, such
problems are chaotic and are compared loosely)."""
import ctypes
import os
import sys
from itertools import chain

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden", "lm.npz")
REF = os.environ.get("SGTA_REFERENCE_ROOT", "/root/reference")
K = np.array([[502.30, 0.0, 319.75], [0.0, 502.30, 179.75], [0.0, 0.0, 1.0]])
N_PROBLEMS = 40


def problems():
    from oracle import lm as olm
    rng = np.random.default_rng(20231)
    for i in range(N_PROBLEMS):
        n = int(rng.integers(4, 8))
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        t = np.array([rng.uniform(-0.3, 0.3), rng.uniform(-0.2, 0.2), rng.uniform(1.0, 2.0)])
        X = rng.uniform(-0.4, 0.4, size=(n, 3))
        P = (olm.rotation_from_quaternion(q) @ X.T).T + t
        uv = (K @ P.T).T
        uv = uv[:, :2] / uv[:, 2:] + rng.normal(0, 0.4, size=(n, 2))
        q0 = q + rng.normal(0, 0.01, size=4)
        q0 /= np.linalg.norm(q0)
        t0 = t + rng.normal(0, 0.01, size=3)
        w = np.array(olm.get_weights_without(n))
        if i % 3 == 0:
            w[:n] = rng.uniform(0.5, 1.0, size=(n, 2))
        yield n, np.hstack([q0, t0]), uv, X, w


def solve_reference_so(so, v0, x2d, x3d, w, n):
    """Exactly the marshalling of register_GN_C (LM.py:256-266)."""
    value_init_l = (ctypes.c_double * 7)(*v0)
    a = (ctypes.c_double * (n * 2))(*chain.from_iterable(x2d.tolist()))
    b = (ctypes.c_double * (n * 3))(*chain.from_iterable(x3d.tolist()))
    wl = (ctypes.c_double * (n * 2 + 2))(*chain.from_iterable(w.tolist()))
    cam = (ctypes.c_double * 9)(*chain.from_iterable(K.tolist()))
    ans = (ctypes.c_double * 7)(*([0.0] * 7))
    so.LM(value_init_l, a, b, wl, cam, ans, n)
    return np.array(list(ans))


def main():
    from oracle import lm as olm
    so = ctypes.cdll.LoadLibrary(os.path.join(REF, "sgtapose", "rf_tools", "libtestso_final.so"))
    out = {"camera": K}
    worst = 0.0
    for i, (n, v0, uv, X, w) in enumerate(problems()):
        ans = solve_reference_so(so, v0, uv, X, w, n)
        mine, its = olm.gn(v0, uv, X, w, K)
        out.update({"n_%d" % i: n, "v0_%d" % i: v0, "x2d_%d" % i: uv, "x3d_%d" % i: X, "w_%d" % i: w,
                    "ans_%d" % i: ans, "its_%d" % i: its})
        d = float(np.abs(ans - mine).max())
        if its < 200:
            worst = max(worst, d)
        print(i, n, "iterations", its, "max |so - restatement| %.2e" % d)
    print("worst over converged problems: %.2e" % worst)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
