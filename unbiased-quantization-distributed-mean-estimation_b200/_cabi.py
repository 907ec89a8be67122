"""ctypes loader of libdme_b200.so (C ABI in include/dme_b200.h).  Fails loudly when the library is missing."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdme_b200.so")

i64, u64, f32p, vp, ci, cf = C.c_int64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_int, C.c_float

# name -> (restype, argtypes); mirrors include/dme_b200.h one to one (tests check the export list against the header)
SIGNATURES = {
    "dme_last_error": (C.c_char_p, []),
    "dme_version": (ci, []),
    "dme_launch_count": (i64, []),
    "dme_profile_enable": (ci, [ci, vp]),
    "dme_profile_read": (ci, [vp, ci]),
    "dme_profile_name": (C.c_char_p, [ci]),
    "dme_set_unbiased_path": (ci, [ci]),
    "dme_set_biased_path": (ci, [ci]),
    "dme_fill_uniforms": (ci, [vp, i64, vp, u64, ci, vp]),
    "dme_add_launches": (None, [i64]),
    "dme_uniform_x": (cf, [u64, u64]),
    "dme_workspace_bytes": (i64, [i64, i64]),
    "dme_codes_bytes": (i64, [i64, i64, i64, ci]),
    "dme_dir_entries": (i64, [i64, i64]),
    "dme_status": (ci, [vp, vp]),
    "dme_l1_norms": (ci, [vp, i64, i64, i64, vp, vp, i64, vp]),
    "dme_type_quantize": (ci, [vp, i64, i64, i64, i64, ci, vp, vp, u64, u64, vp, vp, vp, i64, vp, vp, i64, vp]),
    "dme_type_encode": (ci, [vp, i64, i64, i64, i64, ci, vp, vp, u64, u64, vp, i64, vp, vp, vp, i64, vp]),
    "dme_decode_mean": (ci, [vp, vp, vp, i64, i64, i64, ci, i64, vp, ci, vp]),
    "dme_decode_mean_tiles": (ci, [vp, vp, vp, i64, i64, i64, ci, i64, vp, ci, i64, i64, vp]),
    "dme_quantize_mean": (ci, [vp, i64, i64, i64, i64, ci, vp, u64, u64, i64, vp, ci, vp, i64, vp, vp, vp, i64, vp]),
    "dme_mean_accumulate": (ci, [vp, i64, i64, i64, i64, vp, ci, vp]),
    "dme_peer_sum_slice": (ci, [vp, vp, i64, ci, ci, i64, vp]),
    "dme_hadamard": (ci, [vp, i64, i64, i64, vp]),
    "dme_rht": (ci, [vp, i64, i64, i64, vp, i64, i64, u64, u64, vp, vp]),
    "dme_irht": (ci, [vp, i64, i64, i64, u64, u64, vp, vp]),
    "dme_rademacher": (ci, [vp, i64, u64, vp]),
    "dme_pair_transform": (ci, [vp, i64, i64, i64, vp]),
    "dme_drive": (ci, [vp, i64, i64, i64, vp, i64, u64, vp, ci, vp]),
    "dme_eden_encode": (ci, [vp, i64, i64, i64, i64, ci, u64, u64, vp, vp, vp, vp, vp, vp]),
    "dme_eden_decode": (ci, [vp, vp, i64, i64, i64, ci, u64, u64, vp, vp, vp, i64, vp]),
    "dme_eden_encode_frac": (ci, [vp, i64, i64, i64, i64, ci, ci, cf, vp, u64, u64, vp, vp, vp, vp, vp, vp]),
    "dme_eden_decode_frac": (ci, [vp, vp, i64, i64, i64, ci, ci, cf, vp, vp, cf, u64, u64, vp, vp, vp, i64, vp]),
    "dme_quicfl_encode": (ci, [vp, i64, i64, i64, i64, ci, ci, cf, cf, vp, vp, u64, u64, u64, vp, vp, vp, vp, vp, vp, vp, vp]),
    "dme_quicfl_decode": (ci, [vp, vp, i64, i64, i64, ci, vp, ci, vp, vp, vp, vp, u64, vp, vp, vp, i64, vp]),
    "dme_scalar_quantize": (ci, [vp, i64, i64, i64, cf, u64, u64, vp, vp, i64, vp]),
}

_lib = None


def lib():
    """The loaded library.  Raises (never falls back) when libdme_b200.so has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python __graft_entry__.py build` "
                "(nvcc, sm_100a). There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)          # AttributeError if the .so is stale / incomplete
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def last_error() -> str:
    return (lib().dme_last_error() or b"").decode()
