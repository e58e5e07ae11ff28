"""CPU oracle — NumPy closed-form restatement of the reference hot path.

TEST INFRASTRUCTURE, NOT PRODUCT.  Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s CPU-baseline / `--impl reference` legs may import this package.
The product path (`query_recommendation_system_b200`) never imports it and has
no CPU fallback.

Every function restates (in vectorised NumPy, not copied code) one step of
wamuumu/query-recommendation-system and cites the reference file:line it
follows.  Parity pin: the reference ships no tests or golden vectors
(SURVEY.md §4), so this oracle is pinned against outputs of the UNMODIFIED
reference run in the build container under fixed seeds — see
`tests/golden/make_golden.py` (generator) and `tests/test_oracle_golden.py`
(checker).  Third-party arithmetic on the path: numpy 1.23.1
(`np.random.permutation`, legacy MT19937 stream — version-stable), scikit-learn
1.2.1 `cosine_similarity` (normalise rows, then X·Xᵀ), numba 0.56.4 (plain
sequential NumPy semantics) — requirements.txt:1-5.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import numpy as np

QUERY_WEIGHT = 0.6   # recommender.py:32
USER_WEIGHT = 0.4    # recommender.py:33
DEFAULT_MEAN = 60    # recommender.py:34
LSH_THRESH = 0.2     # recommender.py:153


# --------------------------------------------------------------------------- a1
def permutation_table(m: int, K: int, seed: Optional[int] = None) -> np.ndarray:
    """`K` successive `np.random.permutation(m)` draws from the GLOBAL legacy RNG
    (recommender.py:118-120), returned element-major: `tab[e, i] = perm_i[e]`.

    `seed=None` continues the current global stream (what the reference does).
    """
    if seed is not None:
        np.random.seed(seed)
    tab = np.empty((m, K), dtype=np.int32)
    for i in range(K):
        tab[:, i] = np.random.permutation(m)
    return tab


def signatures(rowptr: np.ndarray, colidx: np.ndarray, perm_tab: np.ndarray) -> np.ndarray:
    """MinHash signature matrix `int64[n, K]` (recommender.py:105-143).

    The reference scans tuples in permuted order and records, for every set, the
    permuted rank of the first tuple that belongs to it (125-129); that is
    `min_{e in set} perm_i[e]`, and sets never touched keep the initial -1 (114).
    """
    n = rowptr.size - 1
    K = perm_tab.shape[1]
    sig = np.full((n, K), -1, dtype=np.int64)
    lens = np.diff(rowptr)
    nz = np.flatnonzero(lens > 0)
    if nz.size == 0:
        return sig
    # process in row chunks to bound the gathered [nnz_chunk, K] temporary
    budget = 1 << 24
    start = 0
    while start < nz.size:
        stop = start
        acc = 0
        while stop < nz.size and (acc == 0 or acc + lens[nz[stop]] * K <= budget):
            acc += lens[nz[stop]] * K
            stop += 1
        rows = nz[start:stop]
        lo, hi = rowptr[rows[0]], rowptr[rows[-1] + 1]
        gathered = perm_tab[colidx[lo:hi]]                      # [nnz_chunk, K]
        offs = (rowptr[rows] - lo).astype(np.int64)
        sig[rows] = np.minimum.reduceat(gathered, offs, axis=0)
        start = stop
    return sig


# ------------------------------------------------------------------------ a2/a3
def band_lanes(sig: np.ndarray, b: int) -> np.ndarray:
    """`int16[n, b, r]` band sub-vectors (lsh.py:17-28): `K % b == 0` is asserted
    (lsh.py:20) and lanes are WRAPPED to int16 (lsh.py:28)."""
    sig = np.asarray(sig)
    n, K = sig.shape
    if K % b != 0:
        raise AssertionError("signature length %d is not a multiple of b=%d" % (K, b))
    return sig.astype(np.int16).reshape(n, b, K // b)


def band_keys_u64(sig: np.ndarray, b: int) -> np.ndarray:
    """Exact bucket ids for r <= 4: the r int16 lanes packed big-endian into one
    u64 (lane 0 most significant).  Two sets share the reference's comma-joined
    string key (lsh.py:32-34) in a band iff they share this integer."""
    lanes = band_lanes(sig, b).astype(np.uint16).astype(np.uint64)
    r = lanes.shape[2]
    if r > 4:
        raise ValueError("64-bit packing is exact only for r <= 4")
    key = np.zeros(lanes.shape[:2], dtype=np.uint64)
    for c in range(r):
        key = (key << np.uint64(16)) | lanes[:, :, c]
    return key


def band_is_skipped(sig: np.ndarray, b: int) -> np.ndarray:
    """`bool[n, b]`: band whose lanes are ALL -1 after the int16 wrap — the
    reference ignores such buckets (lsh.py:44-47), which also swallows genuine
    values congruent to 65535 mod 65536."""
    return (band_lanes(sig, b) == -1).all(axis=2)


def buckets(sig: np.ndarray, b: int):
    """Per band, the bucket partition `{lanes tuple: [set ids in row order]}` —
    the canonical form of `LSH.buckets` after `compute_buckets` has been called
    once per signature in row order (lsh.py:31-38)."""
    lanes = band_lanes(sig, b)
    out = []
    for band in range(b):
        d: Dict[tuple, list] = {}
        for s in range(lanes.shape[0]):
            d.setdefault(tuple(int(v) for v in lanes[s, band]), []).append(s)
        out.append(d)
    return out


# --------------------------------------------------------------------------- a5
def candidates_packed(sig: np.ndarray, b: int) -> np.ndarray:
    """Sorted unique `u64 (i << 32 | j)`, i < j: every pair of sets sharing a
    non-skipped bucket in at least one band (lsh.py:40-54)."""
    lanes = band_lanes(sig, b)
    n = lanes.shape[0]
    skipped = (lanes == -1).all(axis=2)
    chunks = []
    for band in range(b):
        rows = np.flatnonzero(~skipped[:, band])
        if rows.size < 2:
            continue
        sub = lanes[rows, band, :]
        order = np.lexsort(sub.T[::-1])
        srt = sub[order]
        ids = rows[order]
        new = np.ones(ids.size, dtype=bool)
        new[1:] = (srt[1:] != srt[:-1]).any(axis=1)
        starts = np.flatnonzero(new)
        ends = np.append(starts[1:], ids.size)
        for s, e in zip(starts, ends):
            h = e - s
            if h < 2:
                continue
            members = np.sort(ids[s:e]).astype(np.uint64)
            ii, jj = np.triu_indices(h, k=1)
            chunks.append((members[ii] << np.uint64(32)) | members[jj])
    if not chunks:
        return np.zeros(0, dtype=np.uint64)
    return np.unique(np.concatenate(chunks))


def candidates_packed_fast(sig: np.ndarray, b: int) -> np.ndarray:
    """Same result as `candidates_packed` (lsh.py:40-54) without a Python loop per bucket, for the
    full-size parity check of bench.py: band lanes packed into one integer key, one stable sort per
    band, and every member of a bucket paired with the earlier members of its run by index
    arithmetic.  Exact for r <= 4 (the r 16-bit lanes fit one u64); other r fall back."""
    n, K = sig.shape
    r = K // b
    if K % b != 0:
        raise AssertionError("signature length %d is not a multiple of b=%d" % (K, b))
    if r > 4:
        return candidates_packed(sig, b)
    lanes = (sig.astype(np.int64) & 0xFFFF).astype(np.uint64).reshape(n, b, r)
    chunks = []
    for band in range(b):
        key = np.zeros(n, dtype=np.uint64)
        for c in range(r):
            key = (key << np.uint64(16)) | lanes[:, band, c]
        keep = np.flatnonzero(~(lanes[:, band, :] == 0xFFFF).all(axis=1))      # all lanes == -1: skipped (lsh.py:47)
        if keep.size < 2:
            continue
        order = np.argsort(key[keep], kind="stable")
        ids = keep[order]                                  # ascending ids inside a run (stable sort of ascending ids)
        ks = key[keep][order]
        start = np.flatnonzero(np.r_[True, ks[1:] != ks[:-1]])
        run_of = np.cumsum(np.r_[True, ks[1:] != ks[:-1]]) - 1
        off = np.arange(ids.size) - start[run_of]          # earlier members of the same bucket
        tot = int(off.sum())
        if tot == 0:
            continue
        second = np.repeat(np.arange(ids.size), off)       # position of the later member, `off` times
        first_rel = np.arange(tot) - np.repeat(np.cumsum(off) - off, off)
        first = start[run_of[second]] + first_rel
        chunks.append((ids[first].astype(np.uint64) << np.uint64(32)) | ids[second].astype(np.uint64))
    if not chunks:
        return np.zeros(0, dtype=np.uint64)
    return np.unique(np.concatenate(chunks))


def candidates(sig: np.ndarray, b: int) -> set:
    """`set[(i, j)]`, i < j — the return value of `LSH.get_candidates` (lsh.py:40-54)."""
    p = candidates_packed(sig, b)
    return set(zip((p >> np.uint64(32)).astype(np.int64).tolist(),
                   (p & np.uint64(0xFFFFFFFF)).astype(np.int64).tolist()))


# --------------------------------------------------------------------------- a6
def choose_bands(K: int, thresh: float = LSH_THRESH) -> Optional[int]:
    """Band count rule (recommender.py:156-163): the largest `b` dividing `K`
    with `b % 10 == 0` whose `round((1/b)**(1/r), 2) >= thresh`.  Returns None
    where the reference would hit its unbound-`band` error (recommender.py:165)."""
    for b in range(K, 0, -1):
        if K % b == 0 and b % 10 == 0:
            r = K / b
            if round((1 / b) ** (1 / r), 2) >= thresh:
                return b
    return None


def max_candidates(n: int) -> int:
    """`round(log_1.5 n)` (recommender.py:151, 220)."""
    return round(math.log(n, 1.5))


# --------------------------------------------------------------------------- a8
def cosine_rounded(sig: np.ndarray, pairs: np.ndarray) -> np.ndarray:
    """`np.around(cos(sig_i, sig_j), 3)` for `pairs int[P, 2]` (recommender.py:203-204).

    sklearn normalises then multiplies; here the dot and squared norms are exact
    int64 (|value| <= m, K*m^2 < 2^53 for every configured shape) and one fp64
    divide follows, which differs from sklearn by a few ulp and rounds
    identically at 3 decimals except on measure-zero ties."""
    sig = np.asarray(sig, dtype=np.int64)
    pairs = np.asarray(pairs, dtype=np.int64).reshape(-1, 2)
    n2 = (sig * sig).sum(axis=1)
    out = np.empty(pairs.shape[0], dtype=np.float64)
    step = 1 << 16
    for s in range(0, pairs.shape[0], step):
        pi, pj = pairs[s:s + step, 0], pairs[s:s + step, 1]
        dot = (sig[pi] * sig[pj]).sum(axis=1)
        den = np.sqrt(n2[pi].astype(np.float64) * n2[pj].astype(np.float64))
        out[s:s + step] = np.around(dot.astype(np.float64) / den, 3)
    return out


def cosine_milli(sig: np.ndarray, pairs: np.ndarray) -> np.ndarray:
    """The same similarity as the integer `rint(cos * 1000)`; `np.around(x, 3)`
    is `rint(x * 1000) / 1000`, so `cosine_rounded == cosine_milli / 1000.0`."""
    sig = np.asarray(sig, dtype=np.int64)
    pairs = np.asarray(pairs, dtype=np.int64).reshape(-1, 2)
    n2 = (sig * sig).sum(axis=1)
    dot = np.einsum("pk,pk->p", sig[pairs[:, 0]], sig[pairs[:, 1]])
    den = np.sqrt(n2[pairs[:, 0]].astype(np.float64) * n2[pairs[:, 1]].astype(np.float64))
    return np.rint(dot.astype(np.float64) / den * 1000.0).astype(np.int32)


# --------------------------------------------------------------------------- f-3
def centered_rows(rowptr: np.ndarray, val: np.ndarray):
    """Mean-centred rating rows as the reference really builds them
    (recommender.py:263-267): the block is an INTEGER array, so assigning
    `row[rated] - mean` back into it truncates toward zero.  With n rated cells
    summing to S that is the C integer quotient `(r*n - S) / n` (when the exact
    value is an integer, S/n is an integer too and the fp64 subtraction is exact;
    otherwise it is >= 1/n from an integer, far beyond fp64 noise).

    CSR in (ratings uint8, 0 = not rated) -> `cval int16[nnz]`, `n2 int64[rows]`
    (squared norm of each centred row)."""
    rowptr = np.asarray(rowptr, dtype=np.int64)
    val = np.asarray(val).astype(np.int64)
    rows = rowptr.size - 1
    rid = np.repeat(np.arange(rows), np.diff(rowptr))
    rated = val != 0
    cnt = np.bincount(rid[rated], minlength=rows).astype(np.int64)
    tot = np.bincount(rid[rated], weights=val[rated], minlength=rows).astype(np.int64)
    num = val * cnt[rid] - tot[rid]
    den = np.maximum(cnt[rid], 1)
    q = np.abs(num) // den * np.sign(num)                 # truncating division
    cval = np.where(rated, q, 0).astype(np.int16)
    n2 = np.bincount(rid, weights=cval.astype(np.float64) ** 2, minlength=rows).astype(np.int64)
    return cval, n2


def centered_cosine_milli(rowptr, colidx, val, pairs) -> np.ndarray:
    """`rint(1000 * cos)` of the centred rows of each pair, negatives clamped to 0
    (recommender.py:267-269: `np.around(cosine_similarity(block), 3)`, `< 0 -> 0`);
    a row of zero norm has similarity 0 with everything (sklearn's normalize leaves
    it zero).  Exact int64 dot and norms, one fp64 sqrt and divide - the same
    argument as `cosine_milli`."""
    rowptr = np.asarray(rowptr, dtype=np.int64)
    colidx = np.asarray(colidx, dtype=np.int64)
    pairs = np.asarray(pairs, dtype=np.int64).reshape(-1, 2)
    cval, n2 = centered_rows(rowptr, val)
    out = np.zeros(pairs.shape[0], dtype=np.int32)
    for p, (i, j) in enumerate(pairs):
        a0, a1, b0, b1 = rowptr[i], rowptr[i + 1], rowptr[j], rowptr[j + 1]
        _, ia, ib = np.intersect1d(colidx[a0:a1], colidx[b0:b1], assume_unique=True, return_indices=True)
        dot = int((cval[a0:a1][ia].astype(np.int64) * cval[b0:b1][ib].astype(np.int64)).sum())
        if n2[i] == 0 or n2[j] == 0:
            continue
        cos = float(dot) / math.sqrt(float(n2[i]) * float(n2[j]))
        out[p] = max(0, int(np.rint(cos * 1000.0)))
    return out


# --------------------------------------------------------------------------- a7/a9
def topk_tables(n: int, pairs: np.ndarray, sims: np.ndarray, k: int):
    """Neighbour tables (recommender.py:183-210) in CANONICAL order: per set, the
    `k` neighbours with the largest rounded similarity, ties broken by smaller
    neighbour id (the reference's own tie order is CPython-set/argsort dependent
    and unspecified — SURVEY.md appendix).

    Returns dense `idx int32[n, k]` (-1 padded), `val float64[n, k]` (0 padded),
    `cnt int32[n]`.  Sets without any candidate keep `cnt == 0`, matching the
    reference's `if j in querySimilarities` test (recommender.py:314)."""
    pairs = np.asarray(pairs, dtype=np.int64).reshape(-1, 2)
    sims = np.asarray(sims, dtype=np.float64)
    idx = np.full((n, k), -1, dtype=np.int32)
    val = np.zeros((n, k), dtype=np.float64)
    cnt = np.zeros(n, dtype=np.int32)
    if pairs.shape[0] == 0:
        return idx, val, cnt
    src = np.concatenate([pairs[:, 0], pairs[:, 1]])
    dst = np.concatenate([pairs[:, 1], pairs[:, 0]])
    w = np.concatenate([sims, sims])
    order = np.lexsort((dst, -w, src))
    src, dst, w = src[order], dst[order], w[order]
    starts = np.flatnonzero(np.r_[True, src[1:] != src[:-1]])
    ends = np.r_[starts[1:], src.size]
    for s, e in zip(starts, ends):
        c = min(k, e - s)
        row = src[s]
        idx[row, :c] = dst[s:s + c]
        val[row, :c] = w[s:s + c]
        cnt[row] = c
    return idx, val, cnt


# --------------------------------------------------------------------------- a10
def weighted_average(ratings_vec: np.ndarray, indexes: np.ndarray, sims: np.ndarray) -> float:
    """Similarity-weighted mean (recommender.py:36-47).  numba's `np.sum` is a
    plain left-to-right accumulation, reproduced here so the fp64 result is
    bit-identical: num = sum(r_t * s_t), W = sum(s_t for r_t != 0)."""
    num = 0.0
    wsum = 0.0
    for t in range(len(indexes)):
        r = int(ratings_vec[int(indexes[t])])
        s = float(sims[t])
        num = num + float(r) * s
        if r != 0:
            wsum = wsum + s
    if wsum != 0:
        return num / wsum
    return 0.0


def hybrid(q: float, u: float) -> int:
    """Hybrid rule with Python's half-even `round` (recommender.py:324-331)."""
    if u == 0 and q == 0:
        return 0
    if u == 0:
        return round(q * (QUERY_WEIGHT + (USER_WEIGHT * 0.5)) + DEFAULT_MEAN * (USER_WEIGHT * 0.5))
    if q == 0:
        return round(u * (USER_WEIGHT + (QUERY_WEIGHT * 0.5)) + DEFAULT_MEAN * (QUERY_WEIGHT * 0.5))
    return round(q * QUERY_WEIGHT + u * USER_WEIGHT)


# --------------------------------------------------------------------------- a11
def predict_dense(R: np.ndarray,
                  q_tab: Optional[Tuple[np.ndarray, np.ndarray, np.ndarray]],
                  u_tab: Optional[Tuple[np.ndarray, np.ndarray, np.ndarray]]) -> np.ndarray:
    """Dense predictions `int64[U, Q]` (recommender.py:292-343): rated cells keep
    their rating, every zero cell gets the hybrid of the query branch (row `i`
    gathered at query `j`'s neighbours, 314-317) and the user branch (column `j`
    gathered at user `i`'s neighbours, 320).  A table is `(idx, val, cnt)` as
    returned by `topk_tables`, or None for "no such branch" (prediction 0)."""
    R = np.asarray(R)
    U, Q = R.shape
    out = R.astype(np.int64).copy()
    for i, j in np.argwhere(R == 0):
        q = 0.0
        if q_tab is not None and q_tab[2][j] > 0:
            c = q_tab[2][j]
            q = weighted_average(R[i], q_tab[0][j, :c], q_tab[1][j, :c])
        u = 0.0
        if u_tab is not None:
            c = u_tab[2][i]
            u = weighted_average(R[:, j], u_tab[0][i, :c], u_tab[1][i, :c])
        out[i, j] = hybrid(q, u)
    return out


def predict_user_branch_csr(rowptr, colidx, val, Q, u_idx, u_val, u_cnt, rows=None):
    """Binding-B prediction (user branch only, query branch absent → q = 0):
    for each user in `rows`, sparse predictions over the unrated queries, as
    `(cols int32[], preds uint8[])` per user.  Same arithmetic as
    `weighted_average` + `hybrid`, vectorised per neighbour IN TABLE ORDER so
    every cell's fp64 accumulation order matches the reference's."""
    n = rowptr.size - 1
    rows = range(n) if rows is None else rows
    out = {}
    for i in rows:
        num = np.zeros(Q, dtype=np.float64)
        den = np.zeros(Q, dtype=np.float64)
        for t in range(int(u_cnt[i])):
            v = int(u_idx[i, t])
            s = float(u_val[i, t])
            lo, hi = rowptr[v], rowptr[v + 1]
            cj = colidx[lo:hi]
            rv = val[lo:hi].astype(np.float64)
            num[cj] = num[cj] + rv * s
            nzmask = val[lo:hi] != 0
            den[cj[nzmask]] = den[cj[nzmask]] + s
        own = colidx[rowptr[i]:rowptr[i + 1]][val[rowptr[i]:rowptr[i + 1]] != 0]
        den[own] = 0.0
        cols = np.flatnonzero(den != 0)
        u = num[cols] / den[cols]
        x = u * (USER_WEIGHT + (QUERY_WEIGHT * 0.5)) + DEFAULT_MEAN * (QUERY_WEIGHT * 0.5)
        p = np.rint(x)
        p[u == 0] = 0
        keep = p != 0
        out[i] = (cols[keep].astype(np.int32), p[keep].astype(np.int64))
    return out


# --------------------------------------------------------------------------- f-4
def parse_utility_csv(data: bytes):
    """utility_matrix.csv as main.py:69-75 reads it (the file of generator.py:59-77, 329-332: a header of
    query ids without a user field, then `user,r,r,,r,...`; empty = not rated) with Python's own csv module
    (the writer's counterpart) -> (query_ids, user_ids, dense int64[U, Q] with 0 = not rated)."""
    import csv
    import io
    rows = list(csv.reader(io.StringIO(data.decode("utf-8"), newline="")))
    query_ids = rows[0]
    user_ids = [r[0] for r in rows[1:]]
    dense = np.zeros((len(user_ids), len(query_ids)), dtype=np.int64)
    for i, r in enumerate(rows[1:]):
        assert len(r) == len(query_ids) + 1, "row %d: %d fields" % (i, len(r))
        for j, f in enumerate(r[1:]):
            if f != "":
                dense[i, j] = int(f)
    return query_ids, user_ids, dense


def format_matrix_csv(query_ids, user_ids, dense, zero_as_empty: bool) -> bytes:
    """generator.write_csv (generator.py:59-77): `csv.writer(open(..., newline=""))`, header row, then one row
    per user `[id, v, v, ...]`.  zero_as_empty=True is the utility matrix as create_utility_matrix builds it
    (generator.py:329-332: "" = not rated); False is main.py:97-111's final_predictions (plain ints)."""
    import csv
    import io
    buf = io.StringIO(newline="")
    w = csv.writer(buf)
    w.writerow(list(query_ids))
    for u, row in zip(user_ids, np.asarray(dense)):
        w.writerow([u] + [("" if (zero_as_empty and int(v) == 0) else int(v)) for v in row])
    return buf.getvalue().encode("utf-8")
