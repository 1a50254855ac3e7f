"""ctypes binding of libworm_b200.so (include/worm_b200.h).

The shared library is built in-tree by `make -C worm_algorithm_b200/csrc` (or
`__graft_entry__.build()`).  There is no Python or CPU implementation behind it: if the library
is missing, or no sm_100 GPU is visible when a compute entry point is called, this raises.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# WORM_B200_SO: another build of the same library (kernel A/B experiments, tools/build_variant.sh)
SO_PATH = os.environ.get("WORM_B200_SO") or os.path.join(HERE, "libworm_b200.so")

WORM_OK = 0
ALGO_TAYLOR = 0
ALGO_BINARY = 1
ACC_Z, ACC_NB, ACC_NB2, ACC_N, ACC_N2, ACC_ITERS, ACC_STATUS, ACC_STRIDE = 0, 1, 2, 3, 4, 5, 6, 8
MT_Z, MT_NB_TOT, MT_STEPS, MT_NB_FINAL, MT_STATUS, MT_STRIDE = 0, 1, 2, 3, 4, 8


class WormError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__("libworm_b200 error %d: %s" % (code, msg))
        self.code = code


_vp, _i32, _i64, _u64, _sz, _dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_size_t, C.c_double
_pd = C.POINTER(C.c_double)
_pu64 = C.POINTER(C.c_uint64)

# name -> (restype, argtypes); every symbol include/worm_b200.h declares
SIGNATURES = {
    "worm_abi_version": (C.c_int, []),
    "worm_last_error": (C.c_int, [C.c_char_p, _sz]),
    "worm_device_count": (C.c_int, []),
    "worm_device_info": (C.c_int, [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "worm_mt_workspace_bytes": (_sz, [_i64]),
    "worm_mt_run": (C.c_int, [C.c_int, _i64, _pd, _pd, _u64, C.c_int, _vp, _vp, _i32, _vp, _vp, _vp,
                              _i32, _vp, _vp, _vp, _sz, _vp]),
    "worm_philox_workspace_bytes": (_sz, [_i64]),
    "worm_philox_run": (C.c_int, [C.c_int, C.c_int, _i64, _u64, _pd, _pu64, _u64, _u64, _u64,
                                  _vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _u64, _i32, _vp, _vp, _vp,
                                  C.c_int, _vp, _sz, _vp]),
    "worm_thresholds": (C.c_int, [_dbl, _pu64, _pu64, _pu64]),
    "worm_philox_choice": (C.c_int, [C.c_int, C.c_int, _i64, C.c_int, C.POINTER(C.c_int)]),
    "worm_philox_words": (C.c_int, [_u64, _u64, _u64, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "worm_image": (C.c_int, [C.c_int, _i64, _vp, _vp, _vp]),
    "worm_block": (C.c_int, [C.c_int, _i64, _vp, _vp, C.c_int, C.c_int, _vp]),
    "worm_count": (C.c_int, [C.c_int, _i64, _vp, _vp, _vp]),
    "worm_pipeline": (C.c_int, [C.c_int, _i64, _vp, C.c_int, _vp, _vp, _vp]),
    "worm_sim_host": (C.c_int, [C.c_int, _i64, _pd, _pd, _pu64, _u64, _u64, C.c_int, C.c_int, _vp, _vp]),
    "worm_image_host": (C.c_int, [C.c_int, _i64, _vp, _vp]),
    "worm_block_host": (C.c_int, [C.c_int, _i64, _vp, _vp, C.c_int, C.c_int]),
    "worm_pipeline_host": (C.c_int, [C.c_int, _i64, _vp, C.c_int, _vp]),
    "worm_count_host": (C.c_int, [C.c_int, _i64, _vp, _vp]),
    "worm_boundaries": (C.c_int, [C.c_int, C.c_int, _i64, _vp, C.c_int, _vp, _vp, _vp]),
    "worm_boundaries_host": (C.c_int, [C.c_int, C.c_int, _i64, _vp, C.c_int, _vp, _vp]),
    "worm_pca_workspace_bytes": (_sz, [C.c_int, _i64, C.c_int, C.c_int]),
    "worm_pca_timing": (None, [C.c_int]),
    "worm_pca_last_timing": (None, [_pd]),
    "worm_pca": (C.c_int, [C.c_int, _i64, _vp, C.c_int, C.c_int, _dbl, _vp, _vp, _vp,
                           C.POINTER(C.c_int32), _pd, _vp, _sz, _vp]),
    "worm_pca_gram": (C.c_int, [C.c_int, _i64, _vp, C.c_int, _vp, _vp, _sz, _vp]),
    "worm_pca_host": (C.c_int, [C.c_int, _i64, _vp, C.c_int, C.c_int, _dbl, _vp, _vp,
                                C.POINTER(C.c_int32), _pd]),
}

_lib = None


def lib():
    """The loaded library; raises ImportError (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise ImportError(
                "worm_algorithm_b200/libworm_b200.so is missing: build it with "
                "`make -C worm_algorithm_b200/csrc` (nvcc, sm_100a).  There is no CPU fallback.")
        handle = C.CDLL(SO_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)          # AttributeError = ABI mismatch: fail loudly
            fn.restype = res
            fn.argtypes = args
        if handle.worm_abi_version() != 1:
            raise ImportError("libworm_b200.so ABI version %d != 1" % handle.worm_abi_version())
        _lib = handle
    return _lib


def last_error() -> str:
    buf = C.create_string_buffer(512)
    lib().worm_last_error(buf, 512)
    return buf.value.decode("utf-8", "replace")


def check(rc: int) -> None:
    if rc != WORM_OK:
        raise WormError(rc, last_error())


def philox_words(seed: int, chain_id: int, step: int):
    """(a, b) of worm step `step` of chain `chain_id` under key `seed` (stream version 2, include/worm_b200.h)."""
    a, b = C.c_uint32(), C.c_uint32()
    check(lib().worm_philox_words(int(seed), int(chain_id), int(step), C.byref(a), C.byref(b)))
    return a.value, b.value


VARIANT_NAMES = {2: "worm_lat_kernel (32 chains per walker)", 4: "worm_lat_kernel (8 chains per walker)",
                 8: "worm_lat_kernel (4 chains per walker)", 32: "worm_lat_kernel (1 chain per walker)",
                 1: "worm_wide_kernel", -1: "worm_wide_kernel (global memory)", -2: "worm_pack_kernel"}


def philox_choice(L: int, algo: int, n_chains: int, hist: bool = False) -> int:
    """The variant `lanes_per_chain = 0` picks for this batch on the current device."""
    g = C.c_int(0)
    check(lib().worm_philox_choice(int(L), int(algo), int(n_chains), int(bool(hist)), C.byref(g)))
    return g.value


def thresholds(T: float):
    k, t, h = C.c_uint64(), C.c_uint64(), C.c_uint64()
    check(lib().worm_thresholds(float(T), C.byref(k), C.byref(t), C.byref(h)))
    return k.value, t.value, h.value
