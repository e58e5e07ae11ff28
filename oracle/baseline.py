"""CPU baseline: the loop-level port of the reference path (oracle/port.py) timed stage by stage.

TEST / BENCH INFRASTRUCTURE, NOT PRODUCT.  bench.py's `cpu_baseline` leg and `--impl reference`
arm call this on the GPU box, where /root/reference does not exist (kind = "port"); it is the
same per-set / per-cell Python loop structure the reference executes (recommender.py:118-132,
lsh.py:31-54, recommender.py:183-210, 311-331), single-threaded like the reference.
"""
from __future__ import annotations

import time

import numpy as np

from . import port


def run_port_pipeline(rowptr, colidx, val, m, K, b, k, perm_seed=7):
    """Binding B on a closed sample of sets: signatures → LSH candidates → rounded cosine + top-k
    → user-branch predictions for every zero cell.  Returns (timings dict in seconds, counts dict)."""
    n = rowptr.size - 1
    inv = {e: [] for e in range(m)}                      # what compute_shingles hands over (recommender.py:76-95)
    for s in range(n):
        for e in colidx[rowptr[s]:rowptr[s + 1]]:
            inv[int(e)].append(s)
    t = {}
    np.random.seed(perm_seed)
    t0 = time.perf_counter()
    sig = port.signatures_first_touch(inv, m, n, K)
    t["signatures"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    sim, cand = port.query_similarities(sig, b, k)
    t["lsh+cosine+topk"] = time.perf_counter() - t0
    R = np.zeros((n, m), dtype=np.int32)
    for s in range(n):
        R[s, colidx[rowptr[s]:rowptr[s + 1]]] = val[rowptr[s]:rowptr[s + 1]]
    t0 = time.perf_counter()
    cells, final, missed = port.compute_scores(R, {}, sim)
    t["predict"] = time.perf_counter() - t0
    t["total"] = t["signatures"] + t["lsh+cosine+topk"] + t["predict"]
    return t, {"sets": n, "pairs": len(cand), "cells": int(cells.shape[0]), "missed": int(missed.shape[0])}
