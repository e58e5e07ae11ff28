"""CPU baseline, reference arm: the UNMODIFIED reference modules driven over a Binding-B sample.

TEST / BENCH INFRASTRUCTURE, NOT PRODUCT.  `__graft_entry__.build()` stages
`/root/reference/{lsh,recommender}.py` byte for byte into the git-ignored `oracle/_ref/`
(`oracle/stage_ref.py`; the directory ships to the GPU box with the built `.so`, the sources never
enter the history).  This module imports them from there — through the pandas-backed `datatable`
stand-in of `tests/golden/_shim` — and times, per stage, the reference's own functions:

  signatures   `Recommender.compute_signatures()` (recommender.py:105-143) through the
               `compute_shingles` seam (SURVEY.md §8c rule 4: the seam returns `{element: [set ids]}`;
               `dataset.shape[0]`, `queriesIDs.size`, `queries.size` carry the sizes; `PERM` = K)
  LSH          `lsh.LSH(b)`, `compute_buckets` per signature in row order, `get_candidates`
               (lsh.py:12-54; `LSH.buckets` reset first — the class-level list is shared, §8 a4)
  cosine+topk  recommender.py:183-210 lives inside `compute_querySimilarities`, whose band rule
               (156-165) raises for K = 128 / 256 (`band` unbound), so this block alone is the
               loop-level restatement `oracle/port.py:query_similarities` calls — with sklearn's
               own `cosine_similarity`, the reference's arithmetic
  predictions  `Recommender.compute_scores()` (recommender.py:292-343) unmodified: the numba
               `weighted_average` cell loop + hybrid rule over `ratings`, with the two similarity
               builders replaced by closures returning the tables just built (query branch empty,
               user branch = the Binding-B neighbour table)

`available()` is False when `oracle/_ref` was not staged (then bench.py falls back to the port,
kind = "port").
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
SHIM_DIR = os.path.join(os.path.dirname(HERE), "tests", "golden", "_shim")

_mods = None


def available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, "recommender.py")) and os.path.isfile(os.path.join(REF_DIR, "lsh.py"))


def modules():
    """(lsh, recommender) of the reference, imported once from oracle/_ref."""
    global _mods
    if _mods is None:
        if not available():
            raise RuntimeError("oracle/_ref is not staged (run __graft_entry__.build() where /root/reference exists)")
        for p in (SHIM_DIR, REF_DIR):
            if p not in sys.path:
                sys.path.insert(0, p)
        import importlib
        for name in ("lsh", "recommender"):          # a same-named product module must not shadow the reference
            mod = sys.modules.get(name)
            if mod is not None and os.path.dirname(os.path.abspath(getattr(mod, "__file__", ""))) != REF_DIR:
                del sys.modules[name]
        ref_lsh = importlib.import_module("lsh")
        ref_rec = importlib.import_module("recommender")
        assert os.path.dirname(os.path.abspath(ref_lsh.__file__)) == REF_DIR
        assert os.path.dirname(os.path.abspath(ref_rec.__file__)) == REF_DIR
        _mods = (ref_lsh, ref_rec)
    return _mods


class _Shape:
    def __init__(self, rows):
        self.shape = (rows, 1)


def _quiet():
    return contextlib.redirect_stdout(io.StringIO())


def run_reference_pipeline(rowptr, colidx, val, m, K, b, k, perm_seed=7):
    """Binding B on a closed sample of sets, whole path once.  Returns (seconds per stage, counts,
    results) with results = {"sig", "cand", "sim", "final"} for the parity checks of the tests."""
    from . import port
    ref_lsh, ref_rec = modules()
    n = rowptr.size - 1
    inv = {e: [] for e in range(m)}
    for s in range(n):
        for e in colidx[rowptr[s]:rowptr[s + 1]]:
            inv[int(e)].append(s)
    R = np.zeros((n, m), dtype=np.int32)
    for s in range(n):
        R[s, colidx[rowptr[s]:rowptr[s + 1]]] = val[rowptr[s]:rowptr[s + 1]]

    rec = ref_rec.Recommender()
    rec.compute_shingles = lambda: inv
    rec.dataset = _Shape(m)
    rec.queriesIDs = np.arange(n)
    # `queries` is n x features; the early exit at recommender.py:131-132 compares the number of sets seen with
    # queries.size = n * features and therefore never fires in the reference (features >= 2): keep it dead
    rec.queries = np.empty((n, 2), dtype=object)
    t = {}
    old_perm = ref_rec.PERM
    ref_rec.PERM = K
    try:
        np.random.seed(perm_seed)
        t0 = time.perf_counter()
        with _quiet():
            sig = rec.compute_signatures()
        t["signatures"] = time.perf_counter() - t0
    finally:
        ref_rec.PERM = old_perm

    t0 = time.perf_counter()
    ref_lsh.LSH.buckets = []
    lsh = ref_lsh.LSH(b)
    for s in sig:
        lsh.compute_buckets(s)
    cand = lsh.get_candidates(sig)
    t["lsh"] = time.perf_counter() - t0

    t0 = time.perf_counter()
    from sklearn.metrics.pairwise import cosine_similarity
    sim = port.similarities_from_candidates(sig, cand, k, cosine_row0=lambda vecs: cosine_similarity(vecs)[0])
    t["cosine+topk"] = time.perf_counter() - t0

    # the cell loop, unmodified: users = the sample's sets, queries = the elements
    user_sim = {i: sim.get(i, {"indexes": np.zeros(0, dtype=np.int64), "values": np.zeros(0)}) for i in range(n)}
    rec.ratings = R
    rec.usersIDs = np.arange(n)
    rec.queriesIDs = np.arange(m)
    rec.compute_querySimilarities = lambda: {}
    rec.compute_userSimilarities = lambda: user_sim
    t0 = time.perf_counter()
    with _quiet():
        cells, final, missed = rec.compute_scores()
    t["predict"] = time.perf_counter() - t0
    t["lsh+cosine+topk"] = t["lsh"] + t["cosine+topk"]
    t["total"] = t["signatures"] + t["lsh+cosine+topk"] + t["predict"]
    counts = {"sets": n, "pairs": len(cand), "cells": int(cells.shape[0]), "missed": int(missed.shape[0])}
    return t, counts, {"sig": sig, "cand": cand, "sim": sim, "final": final.to_numpy()}


def warm_numba():
    """One tiny call so that numba's JIT of `weighted_average` is not inside a timed step."""
    _, ref_rec = modules()
    ref_rec.weighted_average(np.array([1, 0, 3], dtype=np.int32), np.array([0, 2]), np.array([0.5, 0.25]))
    ref_rec.weighted_average(np.array([1, 0, 3], dtype=np.int32), np.zeros(0, dtype=np.int64), np.zeros(0))
