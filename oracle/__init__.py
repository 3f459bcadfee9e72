"""CPU oracle for the SGTAPose per-frame dense inference path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and there only as the checker (or as
the thing timed on the host cores), never on the CUDA product path.

Every function cites the reference file:line it restates.  Parity pinning:
the reference ships no tests or golden vectors, so the
oracle is pinned against outputs of the reference's own Python, imported from
``/root/reference`` in the build container by ``oracle/make_golden.py`` and
committed under ``tests/golden/``.  The one exception is the DCNv2 operator
itself: its arithmetic lives in the un-vendored third-party extension
lbin/DCNv2 (branch ``pytorch_<ver>``, no pinned commit, README.md:21-28), so at
that single boundary parity is pinned to ``torchvision.ops.deform_conv2d``
(the CPU stand-in BASELINE.json names), not to upstream DCNv2 sources.

Modules: ``dcn`` (deform_conv2d arithmetic), ``model`` (functional forward), ``decode`` (live decode, NMS / top-k,
soft-argmax), ``priors`` and ``detector`` (prior rendering, per-clip host steps), ``preprocess`` (OpenCV's 8-bit
warpAffine + normalisation; pinned to cv2 itself and to the reference's pre_process, ``make_golden_preprocess.py``),
``lm`` (the Gauss-Newton of rf_tools/LM.py; pinned to the outputs of the reference's binary-only
``libtestso_final.so`` and to its Python twin, ``make_golden_lm.py``).  ``ref_import`` and the ``make_golden*``
scripts run in the build container only (they read /root/reference).
"""
