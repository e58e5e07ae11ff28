"""The reference's five manual scenarios (documents/test_descriptions.txt:1-14: data_test 1..5) as property tests
on tiny hand-built utility matrices, run through the CUDA path (C ABI) and — where the oracle has the function —
checked against it as well:

  1. similar queries AND similar users  -> similar results
  2. same users (U1 == U2), different queries -> the same values for both
  3. completely different queries, similar users  -> similar values (user branch alone)
  4. similar queries, completely different users  -> similar values (query branch alone)
  5. completely different queries and users  -> no result

"Similar values" is made exact through the hybrid rule (recommender.py:324-331): with only the user branch a
prediction is round(0.7 u + 18), with only the query branch round(0.8 q + 12), with both round(0.6 q + 0.4 u),
where u / q are similarity-weighted averages of ratings that all lie in a known band.
"""
import numpy as np
import pytest
import torch

from oracle import spec

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _csr(R):
    rows, cols = np.nonzero(R)
    rowptr = np.zeros(R.shape[0] + 1, np.int64)
    rowptr[1:] = np.cumsum((R != 0).sum(axis=1))
    return rowptr, cols.astype(np.int32), R[rows, cols].astype(np.uint8)


def _run_pipeline(R, K=64, b=16, k=8):
    from query_recommendation_system_b200.pipeline import LshCfPipeline
    n, m = R.shape
    rowptr, colidx, val = _csr(R)
    tab = spec.permutation_table(m, K, seed=3)
    pipe = LshCfPipeline(m=m, K=K, b=b, k=k, perm_tab=tab, device=DEV, distributed=False)
    out = pipe.run(torch.from_numpy(rowptr).to(DEV), torch.from_numpy(colidx).to(DEV), torch.from_numpy(val).to(DEV))
    rs, rc = out.pred_row_start.cpu().numpy(), out.pred_row_cnt.cpu().numpy()
    cols, preds = out.pred_cols.cpu().numpy(), out.pred_vals.cpu().numpy()
    P = np.zeros((n, m), np.int64)
    for u in range(n):
        for t in range(rs.shape[1]):
            sl = slice(rs[u, t], rs[u, t] + rc[u, t])
            P[u, cols[sl]] = preds[sl]
    # the oracle on the same tables (the stage-by-stage parity tests pin the tables themselves)
    ref = spec.predict_user_branch_csr(rowptr, colidx, val, m, out.nbr_idx.cpu().numpy(), out.nbr_val.cpu().numpy(),
                                       out.nbr_cnt.cpu().numpy())
    Pref = np.zeros((n, m), np.int64)
    for u, (c, p) in ref.items():
        Pref[u, c] = p
    assert np.array_equal(P, Pref)
    return out, P


def _group(rng, n_users, cols, base, noise, holes):
    """n_users rows that rate `cols` with base[j] +- noise, each leaving `holes` random columns of the group unrated."""
    m = base.size
    R = np.zeros((n_users, m), np.int64)
    for u in range(n_users):
        R[u, cols] = np.clip(base[cols] + rng.integers(-noise, noise + 1, size=cols.size), 1, 100)
        R[u, rng.choice(cols, size=holes, replace=False)] = 0
    return R


def test_scenario_3_similar_users_predict_each_others_values():
    """data_test 3 (and the user half of data_test 1): users of one taste group — same rated queries, ratings within
    +-3 of a common profile — are each other's neighbours, and what a user did not rate is predicted from the group:
    round(0.7 u + 18) with u inside the band of the group's ratings of that query."""
    rng = np.random.default_rng(5)
    m = 120
    base = rng.integers(20, 90, size=m)
    colsA, colsB = np.arange(0, 60), np.arange(60, 120)
    R = np.vstack([_group(rng, 12, colsA, base, 3, 6), _group(rng, 12, colsB, base, 3, 6)])
    out, P = _run_pipeline(R)
    idx, cnt = out.nbr_idx.cpu().numpy(), out.nbr_cnt.cpu().numpy()
    for u in range(24):
        mine = range(0, 12) if u < 12 else range(12, 24)
        assert cnt[u] > 0 and all(int(v) in mine for v in idx[u, :cnt[u]]), "neighbours come from the user's own group"
        cols = colsA if u < 12 else colsB
        for j in cols[R[u, cols] == 0]:
            raters = [v for v in idx[u, :cnt[u]] if R[v, j] != 0]
            if not raters:
                assert P[u, j] == 0
                continue
            lo, hi = min(R[v, j] for v in raters), max(R[v, j] for v in raters)
            assert np.floor(0.7 * lo + 18) <= P[u, j] <= np.ceil(0.7 * hi + 18), (u, j, P[u, j], lo, hi)
        other = colsB if u < 12 else colsA
        assert not P[u, other].any(), "nobody similar rated the other group's queries: missed, not predicted"


def test_scenario_2_identical_users_get_identical_values():
    """data_test 2: U1 == U2 (same rated queries, same ratings) among a larger population — similarity 1.000, each
    is the other's first neighbour, and every jointly unrated query gets the same prediction for both."""
    rng = np.random.default_rng(9)
    m = 200
    base = rng.integers(20, 90, size=m)
    cols = np.arange(0, 80)
    R = _group(rng, 10, cols, base, 4, 8)
    R[1] = R[0]
    R = np.vstack([R, _group(rng, 10, np.arange(100, 200), base, 4, 8)])
    out, P = _run_pipeline(R)
    idx, val = out.nbr_idx.cpu().numpy(), out.nbr_val.cpu().numpy()
    assert idx[0, 0] == 1 and idx[1, 0] == 0 and val[0, 0] == 1.0 and val[1, 0] == 1.0
    # U1 and U2 see the same neighbours except each other; with equal ratings the sums differ only by the twin's
    # own (identical) row, so the predictions coincide wherever both have one
    both = (P[0] != 0) & (P[1] != 0)
    assert both.any() and np.array_equal(P[0][both], P[1][both])
    assert np.array_equal(P[0] != 0, P[1] != 0)


def test_scenario_5_nothing_in_common_gives_no_result():
    """data_test 5: users whose rated queries are pairwise disjoint share no bucket (MinHash values of disjoint sets
    differ in every position), so there are no candidates, no neighbours and every unrated cell is missed."""
    m, n = 400, 16
    R = np.zeros((n, m), np.int64)
    rng = np.random.default_rng(2)
    for u in range(n):
        R[u, u * 25:(u + 1) * 25] = rng.integers(1, 101, size=25)
    out, P = _run_pipeline(R)
    assert out.pairs.numel() == 0 and int(out.nbr_cnt.sum()) == 0 and not P.any() and out.n_predictions == 0


def _dense_predict(R, q_tab, u_tab):
    from query_recommendation_system_b200.recommender import Recommender
    rec = Recommender(device=DEV, verbose=False)
    rec.ratings = R
    return rec.predict_with_tables(q_tab, u_tab)


def test_scenario_4_similar_queries_alone():
    """data_test 4: completely different users (no user table at all), similar queries: an unrated query is predicted
    from the user's OWN ratings of the similar queries — round(0.8 q + 12) (only-query case of the hybrid rule)."""
    rng = np.random.default_rng(4)
    U, Q = 6, 12
    R = rng.integers(30, 80, size=(U, Q)).astype(np.int64)
    R[:, 0] = 0                                   # query 0 is unrated by everybody; queries 1..3 are its neighbours
    q_tab = {0: {"indexes": np.array([1, 2, 3]), "values": np.array([0.9, 0.8, 0.7])}}
    out = _dense_predict(R, q_tab, None)
    for u in range(U):
        q = (0.9 * R[u, 1] + 0.8 * R[u, 2] + 0.7 * R[u, 3]) / (0.9 + 0.8 + 0.7)
        assert abs(out[u, 0] - (0.8 * q + 12)) <= 0.5 + 1e-9
        assert min(R[u, 1:4]) * 0.8 + 12 - 0.5 <= out[u, 0] <= max(R[u, 1:4]) * 0.8 + 12 + 0.5
    assert np.array_equal(out[:, 1:], R[:, 1:])   # rated cells are returned as they are


def test_scenario_1_both_branches_agree_with_their_inputs():
    """data_test 1: similar queries AND similar users: the prediction is round(0.6 q + 0.4 u) and lies between what the
    two branches say on their own."""
    rng = np.random.default_rng(6)
    U, Q = 8, 10
    R = rng.integers(40, 70, size=(U, Q)).astype(np.int64)
    R[0, 0] = 0
    q_tab = {0: {"indexes": np.array([1, 2]), "values": np.array([0.95, 0.9])}}
    u_tab = {0: {"indexes": np.array([1, 2, 3]), "values": np.array([0.9, 0.85, 0.8])}}
    both = _dense_predict(R, q_tab, u_tab)[0, 0]
    q = (0.95 * R[0, 1] + 0.9 * R[0, 2]) / 1.85
    u = (0.9 * R[1, 0] + 0.85 * R[2, 0] + 0.8 * R[3, 0]) / 2.55
    assert abs(both - (0.6 * q + 0.4 * u)) <= 0.5 + 1e-9
    assert min(q, u) - 0.5 <= both <= max(q, u) + 0.5
