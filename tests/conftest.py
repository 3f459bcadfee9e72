import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    import torch
    # fp32 parity: the eager module tree runs its plain convolutions / linears through
    # cuDNN / cuBLAS, which default to TF32 (1e-3 relative per op) unless told otherwise
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "reference: needs /root/reference (build container only)")


def pytest_collection_modifyitems(config, items):
    import torch
    from oracle import ref_import
    has_gpu = torch.cuda.is_available()
    has_ref = ref_import.reference_available()
    has_timeout = config.pluginmanager.hasplugin("timeout")
    for it in items:
        if "gpu" in it.keywords and not has_gpu:
            it.add_marker(pytest.mark.skip(reason="no CUDA device"))
        if "gpu" in it.keywords and has_timeout and it.get_closest_marker("timeout") is None:
            # a deadlocked kernel must fail its test, not hang the run: watchdog thread (works while the main thread
            # sits in a CUDA synchronise), generous limit -- the whole GPU suite takes ~25 s
            it.add_marker(pytest.mark.timeout(600, method="thread"))
        if "reference" in it.keywords and not has_ref:
            it.add_marker(pytest.mark.skip(reason="/root/reference not present"))


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    d = os.path.join(ROOT, "tests", "golden")
    return lambda name: np.load(os.path.join(d, name))
