"""CPU-side checks of the drop-in boundary: libqrs_b200.so loads, exports every symbol that
include/qrs_b200.h declares with the argument count the ctypes table uses, and the argument checks that
need no GPU behave (no kernel is launched here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "qrs_b200.h")


@pytest.fixture(scope="module")
def lib():
    from query_recommendation_system_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return _lib.load()


def header_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(qrs_\w+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        name, args = m.group(1), m.group(2).strip()
        out[name] = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
    return out


def test_header_and_ctypes_table_agree(lib):
    from query_recommendation_system_b200 import _lib
    decl = header_functions()
    assert len(decl) >= 20
    assert set(decl) == set(_lib.SIGNATURES), set(decl) ^ set(_lib.SIGNATURES)
    for name, nargs in decl.items():
        assert hasattr(lib, name), "library does not export %s" % name
        assert len(_lib.SIGNATURES[name][1]) == nargs, name


def test_version_status_strings_and_workspace_queries(lib):
    assert lib.qrs_version() >= 100
    assert lib.qrs_status_string(0) == b"ok" and lib.qrs_status_string(-3) == b"workspace too small"
    assert lib.qrs_sort_workspace_bytes(100_000, 32) > 0
    assert lib.qrs_pair_slots_workspace_bytes(3_200_000) >= 4 * 3_200_000
    assert lib.qrs_topk_workspace_bytes(100_000) > 16 * 100_000
    assert lib.qrs_predict_max_tile(1) % 32 == 0 and lib.qrs_predict_max_tile(0) % 32 == 0
    assert lib.qrs_predict_sparse_workspace_bytes(1000, 50_000, 10_000, 480, 0) >= 1000 * 22 * 8 + 4 * 50_000
    assert lib.qrs_predict_warps_per_sm(1024, 1) >= 16 and lib.qrs_predict_warps_per_sm(1000, 1) == 0
    assert lib.qrs_minhash_max_elements() >= 1_000_000
    # staged peer rows: offset tables for every (user, slot, window) + cap_steps 128-byte steps
    small, big = lib.qrs_peer_rows_stage_bytes(1000, 28, 10_000, 1024, 0), lib.qrs_peer_rows_stage_bytes(1000, 28, 10_000, 1024, 4096)
    assert small >= 1000 * 28 * 10 * 8 and big - small == 4096 * 128
    assert lib.qrs_peer_rows_stage_bytes(1000, 0, 10_000, 1024, 0) == 0


def test_argument_validation_without_a_gpu(lib):
    # invalid arguments are rejected before any CUDA call (QRS_ERR_INVALID_ARG = -1, UNSUPPORTED = -2)
    assert lib.qrs_band_keys(None, 10, 128, 32, None, None, None, None) == -1
    assert lib.qrs_sort_u64(None, None, 10, 1, 0, 64, None, 0, None) == -1
    one = ctypes.c_void_p(8)     # never dereferenced: K % b != 0 is caught first (lsh.py:20)
    assert lib.qrs_band_keys(one, 10, 130, 32, one, None, None, None) == -1
    assert lib.qrs_band_records(one, 10, 128, 300, 0, one, 10, None, None) == -1      # b > QRS_MAX_BANDS
    assert lib.qrs_topk_tables(one, one, 5, 0, 10, 65, one, one, 1 << 20, one, one, None, one, None) == -2   # k > 64
    # staging needs at least two ranks and a 256-byte aligned buffer of the announced size
    assert lib.qrs_peer_rows_stage(one, 1, 0, 100, 1000, 1024, one, one, 8, 100, 0, one, 1 << 20, None) == -1
    assert lib.qrs_peer_rows_stage(one, 2, 0, 100, 1000, 1024, one, one, 8, 100, 0, one, 1 << 20, None) == -3


def test_package_refuses_to_run_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from query_recommendation_system_b200 import LSH, QrsError
    import numpy as np
    lsh = LSH(4)
    lsh.compute_buckets(np.arange(8))
    with pytest.raises(QrsError):
        lsh.get_candidates(None)          # no CPU fallback
