"""CPU-only checks of the host-side logic around the kernels: the reference's parameter rules, the
launch-shape heuristics, the synthetic generator's determinism, and the I/O glue of the Recommender
mirror on the golden C1 files (no kernel is launched)."""
import os

import numpy as np
import pandas as pd
import pytest
import torch

from conftest import GOLDEN


def test_band_rule_matches_reference():
    from query_recommendation_system_b200 import recommender as R
    assert R.choose_bands(180) == 60                      # recommender.py:156-163 at the default PERM
    assert R.choose_bands(128) is None and R.choose_bands(256) is None   # the reference's unbound `band`
    assert (R.PERM, R.QUERY_WEIGHT, R.USER_WEIGHT, R.DEFAULT_MEAN, R.LSH_THRESH) == (180, 0.6, 0.4, 60, 0.2)
    assert R.weighted_average(np.array([0, 50, 100, 0]), np.array([1, 2, 3]), np.array([0.5, 0.25, 0.9])) == \
        (50 * 0.5 + 100 * 0.25) / 0.75                     # recommender.py:36-47: zero ratings carry no weight


def test_launch_heuristics():
    from query_recommendation_system_b200 import stages
    assert stages.choose_head_ranks(10, 3) == 0            # fewer than 32 elements: direct kernel only
    r = stages.choose_head_ranks(10_000, 500)
    assert r % 32 == 0 and 32 <= r <= 2048 and (1 - 0.05) ** r < 1e-6
    assert stages.choose_head_ranks(1_000_000, 1000) == 2048
    from query_recommendation_system_b200 import _lib
    lib = _lib.load()
    max_tile = lib.qrs_predict_max_tile(1)
    assert max_tile % 32 == 0 and 16_000 < max_tile <= 65_536 and lib.qrs_predict_max_tile(0) < max_tile
    top = lib.qrs_predict_warps_per_sm(32, 1)
    for q, avg in ((10_000, 500), (100_000, 1000), (1_000, 50), (37, 3), (1_000_000, 1000)):
        t = stages.choose_predict_tile(q, avg)
        assert t % 32 == 0 and t <= max_tile and -(-q // t) * t >= q
        assert lib.qrs_predict_warps_per_sm(t, 1) == top           # the window never costs a warp
    # the widest window that keeps every warp: 1 024 columns for the integer accumulators, whatever the rows
    assert stages.choose_predict_tile(10_000, 495.0) == 1024                     # C2: ten windows
    assert stages.choose_predict_tile(100_000, 1000.0) in (1024, 992)            # C3: 98 .. 101 windows
    assert lib.qrs_predict_warps_per_sm(1024 + 64, 1) < top
    tf = stages.choose_predict_tile(10_000, 495.0, integer_tables=False)
    assert tf < 1024 and lib.qrs_predict_warps_per_sm(tf, 0) == lib.qrs_predict_warps_per_sm(32, 0)


def test_synth_is_deterministic_and_shardable():
    from query_recommendation_system_b200 import synth
    a = synth.planted_csr(300, 500, 0.05, seed=9)
    b = synth.planted_csr(300, 500, 0.05, seed=9)
    assert all(torch.equal(x, y) for x, y in zip(a, b))
    # any row range regenerates the same rows, also when it starts inside a taste group
    rp, ci, v = synth.planted_csr(100, 500, 0.05, seed=9, row_base=77)
    assert torch.equal(ci, a[1][a[0][77]:a[0][177]]) and torch.equal(v, a[2][a[0][77]:a[0][177]])
    assert int(v.min()) >= 1 and int(v.max()) <= 100       # generator.py:19 vote range
    lens = (a[0][1:] - a[0][:-1]).float()
    assert 15 < lens.mean() < 35                           # ~5 % of 500


def test_recommender_io_glue_on_golden_files(golden_c1):
    from query_recommendation_system_b200.recommender import Recommender
    c1 = os.path.join(GOLDEN, "c1")
    rec = Recommender(verbose=False)
    dataset = pd.read_csv(os.path.join(c1, "dataset.csv"), dtype=str, keep_default_na=False)
    rec.datasetFeatures = list(dataset.columns)[1:]
    queries, ids = rec.parse_queries(os.path.join(c1, "queries.csv"))        # recommender.py:386-416
    assert ids == golden_c1["queries_ids"].tolist() and queries.shape == (100, len(rec.datasetFeatures))
    first = open(os.path.join(c1, "queries.csv")).readline().strip().split(",")
    for cond in first[1:]:
        name, value = cond.split("=")
        assert queries.iloc[0][name] == value
    users = pd.read_csv(os.path.join(c1, "users.csv"), header=None, names=["C0"])
    ratings = pd.read_csv(os.path.join(c1, "utility_matrix.csv"), skiprows=1, header=None, names=["user"] + ids)
    rec.init(users, queries, ids, dataset, ratings)                           # recommender.py:51-64
    assert rec.ratings.dtype == np.int32 and np.array_equal(rec.ratings, golden_c1["ratings"])
    assert rec.dataset.shape[0] == int(golden_c1["m"]) and rec.usersIDs[0] == "U1"


def test_bench_onchip_roofline_arithmetic():
    import bench
    r = bench.onchip_roofline(9.0e8, 6.0, 148, 2000.0)        # 9e8 wavefronts / (148 SMs x 2 GHz) = 3.04 ms
    assert abs(r["floor_ms"] - 9.0e8 / (148 * 2.0e9) * 1e3) < 1e-12 and abs(r["frac"] - r["floor_ms"] / 6.0) < 1e-12
    assert abs(r["achieved"] / r["peak"] - r["frac"]) < 1e-12
