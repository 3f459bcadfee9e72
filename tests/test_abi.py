"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol the
header declares; host-side logic that needs no GPU."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "sgta_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sgta_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from sgtapose_b200 import _lib, build
    build.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = _declared_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), "missing export: " + s
    assert set(syms) == set(_lib.SIGNATURES), "ctypes table and header disagree"
    assert hasattr(lib, "LM"), "the reference's rf_tools entry name (LM.py:10) must be exported too"
    assert lib.sgta_abi_version() == 1


def test_no_cpu_fallback():
    from sgtapose_b200 import _lib
    from sgtapose_b200.dcn_v2 import DCN
    m = DCN(4, 4)
    with pytest.raises(_lib.SgtaError):
        m(torch.zeros(1, 4, 5, 5))


def test_dcn_module_surface():
    from sgtapose_b200.dcn_v2 import DCN
    m = DCN(8, 6, kernel_size=(3, 3), stride=1, padding=1, dilation=1, deformable_groups=1)
    sd = m.state_dict()
    assert list(sd) == ["weight", "bias", "conv_offset_mask.weight", "conv_offset_mask.bias"]
    assert sd["weight"].shape == (6, 8, 3, 3) and sd["conv_offset_mask.weight"].shape == (27, 8, 3, 3)
    assert float(sd["conv_offset_mask.weight"].abs().max()) == 0.0 and float(sd["bias"].abs().max()) == 0.0


def test_last_writer_mask():
    from sgtapose_b200.fusion import _last_writer_mask
    ids = torch.tensor([[5, 3, 5, 7, 3, 3]])
    assert _last_writer_mask(ids).tolist() == [[False, False, True, True, False, True]]
