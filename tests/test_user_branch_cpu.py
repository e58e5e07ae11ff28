"""CPU checks for the f-3 row: the oracle's centred cosine is pinned against the user
neighbour table the UNMODIFIED reference produced on C1 (tests/golden/c1.npz, recommender.py:256-286) and
against the reference's own expression evaluated with sklearn; the pipeline's similarity="ratings" mode is
driven with the oracle-backed stage double."""
import os
import sys

import numpy as np
import torch

from oracle import spec

HERE = os.path.dirname(os.path.abspath(__file__))
if HERE not in sys.path:
    sys.path.insert(0, HERE)


def dense_to_csr(R):
    rows, cols = np.nonzero(R)
    rowptr = np.zeros(R.shape[0] + 1, np.int64)
    np.cumsum(np.bincount(rows, minlength=R.shape[0]), out=rowptr[1:])
    return rowptr, cols.astype(np.int32), R[rows, cols].astype(np.uint8)


def reference_expression(R):
    """recommender.py:263-269 verbatim on one block: integer array, in-place centring, sklearn cosine."""
    from sklearn.metrics.pairwise import cosine_similarity
    block = np.array(R)
    for s in range(len(block)):
        mean = np.mean(block[s][block[s] != 0])
        block[s][block[s] != 0] = block[s][block[s] != 0] - mean
    sim = np.around(cosine_similarity(block), 3)
    np.fill_diagonal(sim, 0)
    sim[sim < 0] = 0
    return block, sim


def test_oracle_matches_reference_user_table_on_c1(golden_c1):
    g = golden_c1
    R = g["ratings"].astype(np.int64)
    rowptr, colidx, val = dense_to_csr(R)
    pairs, want = [], []
    for i in range(R.shape[0]):
        for t in range(int(g["u_cnt"][i])):
            j = int(g["u_idx"][i, t])
            if j != i:                      # the reference lists the user itself (diagonal forced to 0) in small clusters
                pairs.append((i, j))
                want.append(g["u_val"][i, t])
    assert len(pairs) > 700
    got = spec.centered_cosine_milli(rowptr, colidx, val, np.array(pairs))
    assert np.array_equal(got, np.rint(np.array(want) * 1000).astype(np.int32))


def test_oracle_matches_reference_expression_on_random_blocks():
    rng = np.random.default_rng(11)
    for n, q, fill in ((40, 60, 0.3), (25, 200, 0.08), (12, 9, 0.9)):
        R = np.where(rng.random((n, q)) < fill, rng.integers(1, 101, (n, q)), 0).astype(np.int64)
        R[0, :] = 0
        R[0, 0] = 37                         # a single rating centres to 0: zero-norm row
        R[1, :] = np.where(R[1] != 0, 50, 0)  # constant row: zero norm as well
        R[R.sum(axis=1) == 0, 0] = 5          # the reference cannot centre an empty row (mean of nothing)
        block, sim = reference_expression(R)
        rowptr, colidx, val = dense_to_csr(R)
        cval, n2 = spec.centered_rows(rowptr, val)
        assert np.array_equal(cval, block[np.nonzero(R)].astype(np.int16))          # truncation toward zero
        assert np.array_equal(n2, (block * block).sum(axis=1))
        ii, jj = np.triu_indices(n, 1)
        got = spec.centered_cosine_milli(rowptr, colidx, val, np.stack([ii, jj], 1))
        assert np.array_equal(got, np.rint(sim[ii, jj] * 1000).astype(np.int32))
        assert (got == 0).any() and (got > 0).any()        # clamped negatives and positive similarities both occur


def test_pipeline_ratings_mode_with_stage_double():
    import cpu_stage_double as double
    from query_recommendation_system_b200 import synth
    from query_recommendation_system_b200.pipeline import LshCfPipeline
    n, m, K, b, k = 96, 300, 32, 8, 6
    tab = spec.permutation_table(m, K, seed=7)
    rowptr, colidx, val = synth.planted_csr(n, m, 0.08, seed=3)
    pipe = LshCfPipeline(m=m, K=K, b=b, k=k, perm_tab=tab, device="cpu", stages=double, distributed=False)
    pipe.similarity = "ratings"
    out = pipe.run(rowptr, colidx, val, predict=False)
    sig = spec.signatures(rowptr.numpy(), colidx.numpy(), tab)
    cand = np.array(sorted(spec.candidates(sig, b)), dtype=np.int64).reshape(-1, 2)
    assert cand.shape[0] > 50
    p = out.pairs.numpy().view(np.uint64)
    got = {(int(x >> np.uint64(32)), int(x & np.uint64(0xFFFFFFFF))): int(s) for x, s in zip(p, out.sims.numpy())}
    want = spec.centered_cosine_milli(rowptr.numpy(), colidx.numpy(), val.numpy(), cand)
    assert got == {(int(i), int(j)): int(s) for (i, j), s in zip(cand, want)}
    idx, v, cnt = spec.topk_tables(n, cand, want / 1000.0, k)
    assert np.array_equal(out.nbr_idx.numpy(), idx) and np.array_equal(out.nbr_val.numpy(), v)
    assert np.array_equal(out.nbr_cnt.numpy(), cnt)
    assert (want > 300).any()                            # members of one taste group rate alike
