"""Load the UNMODIFIED reference modules from /root/reference (build container only).

The reference cannot be imported as a package (``sgtapose/__init__.py:3-8``
star-imports analysis/datasets -> matplotlib, ruamel.yaml, pyrr ... which are
absent, and ``rf_tools/LM.py:10`` loads a hard-coded absolute path), so the
hot-path files are exec'd by path under their dotted names with an empty
``sgtapose`` package stub (recipe: SURVEY.md appendix C).

The DCN operator is not in the reference tree (``dla.py:21-25`` imports the
third-party ``DCNv2.dcn_v2``); the caller chooses which ``DCN`` class the
reference ``dla.py`` binds: the torchvision stand-in below (oracle) or the
implementation under test.

This module is used only by ``oracle/make_golden.py`` and by CPU tests that
are skipped when /root/reference is absent (it does not exist on the GPU box).
"""
import importlib.util
import math
import os
import sys
import types

import torch
from torch import nn

REF_ROOT = os.environ.get("SGTA_REFERENCE_ROOT", "/root/reference")
REF_PKG = os.path.join(REF_ROOT, "sgtapose")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_PKG, "lib", "model", "networks", "dla.py"))


class TorchvisionDCN(nn.Module):
    """Stand-in for lbin/DCNv2 ``DCN`` (constructed at dla.py:545).

    Same constructor keywords and state-dict names as upstream
    (weight, bias, conv_offset_mask.{weight,bias}); arithmetic is
    ``torchvision.ops.deform_conv2d`` with offset = first 18 channels of
    ``conv_offset_mask(x)`` and mask = sigmoid of the last 9.
    """

    def __init__(self, in_channels, out_channels, kernel_size=(3, 3), stride=1,
                 padding=1, dilation=1, deformable_groups=1):
        super().__init__()
        if isinstance(kernel_size, int):
            kernel_size = (kernel_size, kernel_size)
        self.kernel_size = tuple(kernel_size)
        self.stride, self.padding, self.dilation = stride, padding, dilation
        self.deformable_groups = deformable_groups
        kh, kw = self.kernel_size
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels, kh, kw))
        self.bias = nn.Parameter(torch.zeros(out_channels))
        self.conv_offset_mask = nn.Conv2d(in_channels, deformable_groups * 3 * kh * kw,
                                          kernel_size=self.kernel_size, stride=stride,
                                          padding=padding, bias=True)
        stdv = 1.0 / math.sqrt(in_channels * kh * kw)
        with torch.no_grad():
            self.weight.uniform_(-stdv, stdv)
            self.conv_offset_mask.weight.zero_()
            self.conv_offset_mask.bias.zero_()

    def forward(self, x):
        from torchvision.ops import deform_conv2d
        om = self.conv_offset_mask(x)
        n_off = 2 * self.deformable_groups * self.kernel_size[0] * self.kernel_size[1]
        return deform_conv2d(x, om[:, :n_off], self.weight, self.bias, stride=self.stride,
                             padding=self.padding, dilation=self.dilation,
                             mask=torch.sigmoid(om[:, n_off:]))


def _load(dotted, path):
    spec = importlib.util.spec_from_file_location(dotted, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[dotted] = mod
    spec.loader.exec_module(mod)
    return mod


def load_reference(dcn_cls=None, dcn_shim_dir=None):
    """Return a namespace with the reference's dla, base_model, utils, decode,
    image_proc and spatial_softmax modules, ``dla.DCN`` bound to ``dcn_cls``
    (default: the torchvision stand-in).  With ``dcn_shim_dir`` the reference's own relative import
    (``dla.py:22``) resolves ``DCNv2/dcn_v2.py`` from that directory instead -- the way a maintainer
    installs the shim of ``integration/`` (level 1)."""
    if not reference_available():
        raise RuntimeError("reference tree not found at %s" % REF_ROOT)
    dcn_cls = dcn_cls or TorchvisionDCN
    for name in list(sys.modules):
        if name == "sgtapose" or name.startswith("sgtapose."):
            del sys.modules[name]

    def pkg(name, path):
        m = types.ModuleType(name)
        m.__path__ = [path]
        sys.modules[name] = m
        return m

    root = pkg("sgtapose", REF_PKG)
    pkg("sgtapose.lib", os.path.join(REF_PKG, "lib"))
    pkg("sgtapose.lib.model", os.path.join(REF_PKG, "lib", "model"))
    nets = os.path.join(REF_PKG, "lib", "model", "networks")
    pkg("sgtapose.lib.model.networks", nets)
    if dcn_shim_dir is not None:
        pkg("sgtapose.lib.model.networks.DCNv2", dcn_shim_dir)
    else:
        pkg("sgtapose.lib.model.networks.DCNv2", os.path.join(nets, "DCNv2"))
        dcn_mod = types.ModuleType("sgtapose.lib.model.networks.DCNv2.dcn_v2")
        dcn_mod.DCN = dcn_cls
        sys.modules[dcn_mod.__name__] = dcn_mod
    # image_proc.py:7,13 import matplotlib.pyplot and webcolors, unused on this path
    for stub in ("matplotlib", "matplotlib.pyplot", "webcolors"):
        if stub not in sys.modules:
            try:
                importlib.import_module(stub)
            except Exception:
                sys.modules[stub] = types.ModuleType(stub)
    if not hasattr(sys.modules["matplotlib"], "pyplot"):
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]

    ns = types.SimpleNamespace()
    ns.base_model = _load("sgtapose.lib.model.networks.base_model", os.path.join(nets, "base_model.py"))
    ns.dla = _load("sgtapose.lib.model.networks.dla", os.path.join(nets, "dla.py"))
    ns.image_proc = _load("sgtapose.image_proc", os.path.join(REF_PKG, "image_proc.py"))
    root.image_proc = ns.image_proc
    ns.utils = _load("sgtapose.lib.model.utils", os.path.join(REF_PKG, "lib", "model", "utils.py"))
    ns.decode = _load("sgtapose.lib.model.decode", os.path.join(REF_PKG, "lib", "model", "decode.py"))
    ns.spatial_softmax = _load("sgtapose.spatial_softmax", os.path.join(REF_PKG, "spatial_softmax.py"))
    return ns


from sgtapose_b200.config import HEAD_CONV, HEADS, default_opt  # noqa: E402,F401  (shared description)


def build_reference_model(ns=None, opt=None):
    import contextlib
    import io
    ns = ns or load_reference()
    opt = opt or default_opt()
    with contextlib.redirect_stdout(io.StringIO()):
        model = ns.dla.DLA_PlanAWindow_l3new(34, dict(HEADS), dict(HEAD_CONV), opt).eval()
    return model
