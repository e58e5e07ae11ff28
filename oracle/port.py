"""CPU oracle — loop-level Python port of the reference hot path.

TEST INFRASTRUCTURE, NOT PRODUCT (see `oracle/spec.py` header for the rules).
Where `spec.py` gives closed forms, this module keeps the reference's own data
structures and iteration order (string bucket keys, one global Python `set` of
pairs, per-cell Python loop), so that (a) it reproduces the reference's
unspecified tie order on the same interpreter and (b) timing it is a fair
stand-in for "the reference's CPU path" on a box where `/root/reference` does
not exist (`bench.py` `cpu_baseline.kind == "port"`).  It is written from the
behaviour documented in SURVEY.md §8(a), not copied; each function cites the
lines it follows.  Pinned against the live reference by
`tests/golden/make_golden.py` → `tests/test_oracle_golden.py`.
"""
from __future__ import annotations

import math
from itertools import combinations

import numpy as np

from . import spec


def signatures_first_touch(inverted, m, n, K):
    """recommender.py:114-139 — for each of `K` permutations drawn from the
    global RNG, walk the elements by increasing permuted rank and give every set
    the rank of the first element that contains it.  `inverted[e]` lists the set
    ids containing element `e` (what `compute_shingles` returns, 76-95)."""
    sig = np.full((K, n), -1, dtype=int)
    for h in range(K):
        ranks = np.random.permutation(m)
        walk = np.argsort(ranks)
        seen = set()
        for e in walk:
            for s in inverted[e]:
                if s not in seen:
                    seen.add(s)
                    sig[h][s] = ranks[e]
        # the reference's early exit compares against n*features and never fires (131-132)
    return sig.T


class LSHPort:
    """lsh.py:7-54 with instance-scoped state (the reference's class-level
    `buckets` list is a sharing bug)."""

    def __init__(self, b):
        self.b = b
        self.tables = [dict() for _ in range(b)]
        self.next_id = 0

    def make_subvecs(self, signature):
        K = len(signature)
        assert K % self.b == 0
        r = K // self.b
        return np.stack([signature[o:o + r] for o in range(0, K, r)]).astype("int16")

    def compute_buckets(self, signature):
        for band, lanes in enumerate(self.make_subvecs(signature).astype(str)):
            self.tables[band].setdefault(",".join(lanes), []).append(self.next_id)
        self.next_id += 1

    def get_candidates(self, signatures=None):
        found = set()
        for table in self.tables:
            for key, members in table.items():
                if len(members) > 1 and set(key.split(",")) != {"-1"}:
                    for pair in list(combinations(members, 2)):
                        found.add(pair)
        return found


def _cosine_row0(vectors):
    """sklearn `cosine_similarity(X)[0]` (recommender.py:204): L2-normalise the
    rows in fp64, then one matrix product."""
    X = np.asarray(vectors, dtype=np.float64)
    norms = np.sqrt((X * X).sum(axis=1))
    norms[norms == 0.0] = 1.0
    Xn = X / norms[:, None]
    return Xn @ Xn[0]


def query_similarities(signatures, b, k):
    """recommender.py:171-210 — LSH candidates → adjacency in set-iteration
    order → rounded cosine of each set against its neighbours → top-k by
    `argsort()[::-1]`.  Returns `{i: {"indexes": int array, "values": fp64 array}}`
    containing only sets that have at least one candidate."""
    lsh = LSHPort(b)
    for s in signatures:
        lsh.compute_buckets(s)
    cand = lsh.get_candidates(signatures)
    return similarities_from_candidates(signatures, cand, k), cand


def similarities_from_candidates(signatures, cand, k, cosine_row0=_cosine_row0):
    """recommender.py:183-210 alone (the block after `get_candidates`): adjacency in the
    candidate set's iteration order, rounded cosine of each set against its neighbours,
    top-k by `argsort()[::-1]`.  `cosine_row0(vectors)` = row 0 of the cosine matrix."""
    sim = {}
    for i, j in cand:
        for a in (i, j):
            if a not in sim:
                sim[a] = {"indexes": [], "signatures": [signatures[a]]}
        sim[i]["indexes"].append(j)
        sim[j]["indexes"].append(i)
        sim[i]["signatures"].append(signatures[j])
        sim[j]["signatures"].append(signatures[i])
    for i in sim:
        vals = np.around(cosine_row0(sim[i].pop("signatures"))[1:], 3)
        keep = np.argsort(vals)[::-1][:k]
        sim[i]["indexes"] = np.array(sim[i]["indexes"])[keep]
        sim[i]["values"] = vals[keep]
    return sim


def weighted_average(ratings_vec, indexes, sims):
    """recommender.py:36-47 with numba's left-to-right `np.sum`."""
    return spec.weighted_average(ratings_vec, indexes, sims)


def compute_scores(ratings, query_sim, user_sim):
    """recommender.py:301-341 — per zero cell: query branch if the query has a
    neighbour entry, user branch always, hybrid rule, Python `round`."""
    final = ratings.copy()
    cells = np.array(np.where(ratings == 0)).T
    ratings_t = ratings.T
    for i, j in cells:
        q = weighted_average(ratings[i], query_sim[j]["indexes"], query_sim[j]["values"]) if j in query_sim else 0
        u = weighted_average(ratings_t[j], user_sim[i]["indexes"], user_sim[i]["values"]) if i in user_sim else 0
        final[i][j] = spec.hybrid(q, u)
    missed = np.array(np.where(final == 0)).T
    return cells, final, missed


def dict_to_table(sim, n, k):
    """Neighbour dict → dense `(idx int32[n,k], val fp64[n,k], cnt int32[n])`,
    keeping the dict's own (reference) neighbour order."""
    idx = np.full((n, k), -1, dtype=np.int32)
    val = np.zeros((n, k), dtype=np.float64)
    cnt = np.zeros(n, dtype=np.int32)
    for i, e in sim.items():
        c = len(e["indexes"])
        assert c <= k
        idx[i, :c] = e["indexes"]
        val[i, :c] = e["values"]
        cnt[i] = c
    return idx, val, cnt


def max_candidates(n):
    return round(math.log(n, 1.5))
