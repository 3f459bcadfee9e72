"""ctypes binding of libsgta_b200.so (the C ABI declared in include/sgta_b200.h).

The product path has NO CPU fallback: if the shared library is missing or a kernel call
fails, this module raises.  PyTorch is used only for device memory and streams.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsgta_b200.so")

_c = ctypes
_P = _c.c_void_p
_I = _c.c_int
_F = _c.c_float
_L = _c.c_int64



class SgtaPlanes(ctypes.Structure):
    """struct sgta_planes (include/sgta_b200.h)."""
    _fields_ = [("data", _c.c_void_p), ("rows", _c.c_int64), ("guard", _c.c_int32), ("nchunks", _c.c_int32),
                ("chunk0", _c.c_int32), ("nplanes", _c.c_int32), ("layout", _c.c_int32), ("border", _c.c_int32),
                ("B", _c.c_int32), ("H", _c.c_int32), ("W", _c.c_int32)]


_V = _c.POINTER(SgtaPlanes)

# name -> (restype, argtypes); mirrors include/sgta_b200.h one to one
SIGNATURES = {
    "sgta_debug_flags": (_I, [_I]),
    "sgta_planes_guard": (_I, [_I]),
    "sgta_planes_ntile": (_I, [_I, _I]),
    "sgta_planes_wpack_bytes": (_c.c_int64, [_I, _I, _I]),
    "sgta_planes_pack_weight": (_I, [_P, _P, _I, _I, _I, _P]),
    "sgta_planes_conv": (_I, [_V, _P, _P, _P, _V, _V, _P, _L] + [_I] * 7 + [_P]),
    "sgta_planes_conv_heads": (_I, [_V, _P, _P, _P, _P, _P, _c.POINTER(_P), _c.POINTER(_I), _I, _I, _I, _I, _P]),
    "sgta_planes_conv_sc": (_I, [_V, _P, _P, _P, _V] + [_I] * 7 + [_c.POINTER(_I), _I, _I, _P]),
    "sgta_planes_dcn": (_I, [_V, _P, _P, _P, _P, _V, _I, _I, _I, _P]),
    "sgta_planes_from_nchw": (_I, [_P, _V, _I, _I, _P]),
    "sgta_planes_to_nchw": (_I, [_V, _P, _I, _I, _P]),
    "sgta_planes_pack_stem": (_I, [_P, _P, _V, _I, _I, _P]),
    "sgta_planes_maxpool2": (_I, [_V, _I, _V, _I, _I, _P]),
    "sgta_planes_upsample_add": (_I, [_V, _P, _V, _V, _I, _I, _P]),
    "sgta_planes_superpixels": (_I, [_V, _V, _I, _P]),
    "sgta_planes_gather_tokens": (_I, [_V, _I, _P, _P, _I, _I, _I, _P]),
    "sgta_planes_scatter_tokens": (_I, [_V, _I, _P, _P, _I, _I, _I, _P]),
    "sgta_abi_version": (_I, []),
    "sgta_last_error": (_c.c_char_p, []),
    "sgta_launch_count": (_c.c_int64, []),
    "sgta_dcn_forward": (_I, [_P] * 5 + [_I] * 12 + [_P]),
    "sgta_dcn_backward": (_I, [_P] * 8 + [_I] * 12 + [_P]),
    "sgta_attn_forward": (_I, [_P] * 5 + [_I] * 5 + [_F, _P]),
    "sgta_attn_backward": (_I, [_P] * 9 + [_I] * 5 + [_F, _P]),
    "sgta_topk_index": (_I, [_P, _P, _I, _I, _I, _I, _P]),
    "sgta_window_ids": (_I, [_P, _P, _I, _I, _I, _F, _I, _I, _I, _P]),
    "sgta_gather_tokens": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "sgta_scatter_tokens": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "sgta_decode_peaks": (_I, [_P] * 9 + [_c.POINTER(_c.c_double)] + [_I] * 4 + [_P]),
    "sgta_decode_peaks_exact64": (_I, [_P] * 9 + [_c.POINTER(_c.c_double)] + [_I] * 4 + [_P]),
    "sgta_decode_recheck_count": (_I, [_c.POINTER(_c.c_uint64), _I]),
    "sgta_decode_full_map": (_I, [_I]),
    "sgta_decode_nms_topk": (_I, [_P] * 5 + [_I] * 5 + [_P]),
    "sgta_nms3x3": (_I, [_P, _P, _I, _I, _I, _I, _P]),
    "sgta_soft_argmax": (_I, [_P, _P, _I, _I, _I, _I, _F, _F, _P]),
    "sgta_token_mlp": (_I, [_P] * 15 + [_I] * 4 + [_F, _P]),
    "sgta_token_linear": (_I, [_P, _I, _P, _I, _P, _P, _P, _I, _I, _I, _P]),
    "sgta_token_linear_heads": (_I, [_P, _I, _P, _P, _I, _I, _I, _I, _P]),
    "sgta_attn_kvhm_supported": (_I, [_I] * 6),
    "sgta_attn_forward_kvhm": (_I, [_P] * 5 + [_I] * 5 + [_F, _P]),
    "sgta_preprocess": (_I, [_P, _P, _P, _c.POINTER(_c.c_double), _I, _c.POINTER(_F), _c.POINTER(_F)] + [_I] * 5 + [_P]),
    "sgta_post_process": (_I, [_P, _P, _P, _c.POINTER(_F), _F, _c.c_double, _I, _I, _P]),
    "sgta_lm_refine": (_I, [_c.POINTER(_c.c_double)] * 6 + [_I]),
    "sgta_render_priors": (_I, [_P, _P, _P, _P, _c.POINTER(_F)] + [_I] * 6 + [_P]),
}

_lib = None


class SgtaError(RuntimeError):
    pass


def load():
    """Load (once) and return the ctypes handle; raises if the extension is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SgtaError(
            "libsgta_b200.so not found at %s -- build it with `python -m sgtapose_b200.build` "
            "(there is no CPU fallback)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def call(name, *args):
    """Invoke an int-returning entry point; raise SgtaError with the library's message."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise SgtaError("%s failed (%d): %s" % (name, rc, lib.sgta_last_error().decode()))


def launch_count():
    return int(load().sgta_launch_count())


def ptr(t):
    """Device pointer of a contiguous CUDA tensor (or None)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise SgtaError("expected a CUDA tensor (the B200 path has no CPU fallback)")
    if not t.is_contiguous():
        raise SgtaError("expected a contiguous tensor")
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream
