"""ctypes binding of libqrs_b200.so (the C ABI declared in include/qrs_b200.h).

There is NO CPU fallback: importing this module without the built shared library,
or calling any stage without a CUDA device, raises.  Build the library with
`python __graft_entry__.py` (or `make -C query_recommendation_system_b200/csrc`).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("QRS_B200_LIB", os.path.join(_HERE, "libqrs_b200.so"))   # override: debugging builds only

QRS_OK = 0
MINHASH_AUTO, MINHASH_DIRECT, MINHASH_FIRST_TOUCH, MINHASH_TILE = 0, 1, 2, 3
PAIRS_GROUPED, PAIRS_FORCE_SEARCH = 1, 2
MAX_TOPK = 64
MAX_BANDS = 256
HASH_BITS = 24
SENTINEL_HASH = 0xFFFFFF


class QrsError(RuntimeError):
    pass


_p = C.c_void_p
_i32, _i64, _sz = C.c_int32, C.c_int64, C.c_size_t

# name -> (restype, argtypes); mirrors include/qrs_b200.h one to one
SIGNATURES = {
    "qrs_version": (C.c_int, []),
    "qrs_status_string": (C.c_char_p, [C.c_int]),
    "qrs_last_cuda_error": (C.c_char_p, []),
    "qrs_device_sm_count": (C.c_int, [C.c_int]),
    "qrs_launch_count": (C.c_ulonglong, []),
    "qrs_perm_head_elems": (_i64, [_i32, _i32]),
    "qrs_perm_head_invert": (C.c_int, [_p, _i32, _i32, _i32, _p, _p]),
    "qrs_minhash_max_elements": (_i64, []),
    "qrs_minhash_perm": (C.c_int, [_p, _p, _i64, _i32, _p, _p, _i32, _i32, _i32, _i64, _p, _p, _p, _i64, _i32, _p]),
    "qrs_band_keys": (C.c_int, [_p, _i64, _i32, _i32, _p, _p, _p, _p]),
    "qrs_band_records": (C.c_int, [_p, _i64, _i32, _i32, _i64, _p, _i64, _p, _p]),
    "qrs_sort_workspace_bytes": (_sz, [_i64, _i32]),
    "qrs_sort_u64": (C.c_int, [_p, _p, _i64, _i32, _i32, _i32, _p, _sz, _p]),
    "qrs_pair_slots_workspace_bytes": (_sz, [_i64]),
    "qrs_pair_slots": (C.c_int, [_p, _i64, _p, _p, _sz, _p]),
    "qrs_pairs_score": (C.c_int, [_p, _i64, _p, _p, _p, _i64, _i32, _i32, _p, _p, _i64, _p, _p]),
    "qrs_pair_cosine": (C.c_int, [_p, _p, _i32, _p, _i64, _p, _p]),
    "qrs_topk_workspace_bytes": (_sz, [_i64]),
    "qrs_topk_tables": (C.c_int, [_p, _p, _i64, _i64, _i64, _i32, _p, _p, _sz, _p, _p, _p, _p, _p]),
    "qrs_answer_sets_workspace_bytes": (_sz, [_i64]),
    "qrs_answer_sets_count": (C.c_int, [_p, _i64, _i32, _p, _i64, _p, _p, _sz, _p]),
    "qrs_answer_sets_fill": (C.c_int, [_p, _i64, _i32, _p, _i64, _p, _p, _p]),
    "qrs_recommend_topn": (C.c_int, [_p, _p, _p, _i32, _p, _i64, _i32, _i32, _p, _p, _p, _p]),
    "qrs_peer_alloc": (C.c_int, [_sz, C.POINTER(C.c_void_p)]),
    "qrs_peer_free": (C.c_int, [_p]),
    "qrs_peer_export": (C.c_int, [_p, _p]),
    "qrs_peer_open": (C.c_int, [_p, C.POINTER(C.c_void_p)]),
    "qrs_peer_close": (C.c_int, [_p]),
    "qrs_group_pad": (C.c_int, [_p, _i64, _i32, _i64, _p, _p, _p, _p]),
    "qrs_route_pairs": (C.c_int, [_p, _p, _p, _i64, _i64, _i32, _i64, _p, _p, _p, _p]),
    "qrs_unpack_routed": (C.c_int, [_p, _i64, _p, _p, _p]),
    "qrs_csv_workspace_bytes": (_sz, [_i64, _i64]),
    "qrs_csv_line_ends": (C.c_int, [_p, _i64, _p, _i64, _p, _p, _sz, _p]),
    "qrs_csv_ratings_count": (C.c_int, [_p, _i64, _p, _i64, _i64, _i32, _p, _p, _p, _p, _sz, _p]),
    "qrs_csv_ratings_fill": (C.c_int, [_p, _i64, _p, _i64, _i64, _i32, _p, _p, _p, _p]),
    "qrs_csv_format_workspace_bytes": (_sz, [_i64]),
    "qrs_csv_format_count": (C.c_int, [_p, _p, _p, _p, _p, _p, _i32, _i64, _i32, _i32, _p, _p, _sz, _p]),
    "qrs_cols_widen_u16": (C.c_int, [_p, _i64, _p, _p]),
    "qrs_csv_format_fill": (C.c_int, [_p, _p, _p, _p, _p, _p, _p, _p, _i32, _p, _i32, _i64, _i32, _i32, _p, _p, _p]),
    "qrs_center_rows": (C.c_int, [_p, _p, _i64, _p, _p, _p]),
    "qrs_pair_centered_cosine": (C.c_int, [_p, _p, _p, _p, _p, _p, _i64, _i32, _i32, _p, _p]),
    "qrs_predict_dense": (C.c_int, [_p, _p, _p, _i64, _i32, _p, _p, _p, _i32, _p, _p, _p, _i32, _i64, _i64, _p, _p]),
    "qrs_predict_max_tile": (_i32, [_i32]),
    "qrs_predict_warps_per_sm": (_i32, [_i32, _i32]),
    "qrs_predict_sparse_workspace_bytes": (_sz, [_i64, _i64, _i32, _i32, _i64]),
    "qrs_predict_prepare": (C.c_int, [_p, _p, _p, _i64, _i64, _i32, _i32, _i64, _p, _sz, _p]),
    "qrs_peer_rows_stage_bytes": (_sz, [_i64, _i32, _i32, _i32, _i64]),
    "qrs_peer_rows_stage": (C.c_int, [_p, _i32, _i32, _i64, _i32, _i32, _p, _p, _i32, _i64, _i64, _p, _sz, _p]),
    "qrs_predict_sparse": (C.c_int, [_p, _p, _p, _i64, _i64, _i32, _p, _p, _p, _p, _i32, _i64, _i64, _i32, _p, _sz, _i32,
                                     _p, _i32, _i32, _p, _i64, _p, _p, _p, _i32, _p, _i64, _i32, _p, _p]),
}

_lib = None


def load():
    """Load the shared library once; raise (never fall back) when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise QrsError(
            "libqrs_b200.so is not built (%s). Run `python __graft_entry__.py` or "
            "`make -C query_recommendation_system_b200/csrc`; there is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here == header/library drift
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int, what: str):
    if status == QRS_OK:
        return
    lib = load()
    msg = lib.qrs_status_string(status).decode()
    if status == -4:
        msg += ": " + lib.qrs_last_cuda_error().decode()
    raise QrsError("%s failed: %s" % (what, msg))


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else t.data_ptr()
