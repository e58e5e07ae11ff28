#!/usr/bin/env python
"""Generate the golden vectors by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference, which does not exist on
the GPU box):   python tests/golden/make_golden.py

The reference has no tests or golden vectors of its own, so the
pins are outputs of the reference itself under fixed seeds:

  c1.npz + c1/*.csv   resources/generator.py defaults (random.seed(1234)) pushed
                      through main.py's loading sequence and
                      Recommender.compute_scores() with np.random.seed(7):
                      answer sets, signature matrix, LSH buckets, candidate
                      set, both neighbour tables, predictions, missed cells.
  synth_*.npz         seeded synthetic sets × elements CSR injected through the
                      compute_shingles seam:
                      Recommender.compute_signatures() + LSH(b) buckets and
                      candidates + the cosine/top-k block, for several (K, b),
                      including empty sets, r = 8, and m > 65536 (int16 wrap).
  wavg.npz            weighted_average (numba) + hybrid-rule cases incl. .5 ties.

Harness rules followed: np.random.seed immediately before
compute_signatures; lsh.LSH.buckets reset before each LSH(b); random re-seeded
after `import generator`; generator run with cwd = a scratch copy of resources/.
"""
import contextlib
import io
import os
import random
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
sys.path.insert(0, os.path.join(HERE, "_shim"))     # datatable stand-in
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(REF, "resources"))

import lsh as ref_lsh                   # noqa: E402  (reference, unmodified)
import recommender as ref_rec           # noqa: E402
import generator as ref_gen             # noqa: E402
from datatable import dt                # noqa: E402  (shim)


def quiet():
    return contextlib.redirect_stdout(io.StringIO())


def buckets_to_arrays(tables, r):
    """LSH.buckets (list of {str key: [ids]}) → flat arrays."""
    band, lanes, offs, members = [], [], [0], []
    for bi, table in enumerate(tables):
        for key, ids in table.items():
            band.append(bi)
            lanes.append([int(x) for x in key.split(",")])
            members.extend(ids)
            offs.append(len(members))
    return (np.array(band, np.int32), np.array(lanes, np.int16).reshape(-1, r),
            np.array(offs, np.int64), np.array(members, np.int32))


def sim_to_arrays(sim, n):
    kmax = max([len(v["indexes"]) for v in sim.values()] + [1])
    idx = np.full((n, kmax), -1, np.int32)
    val = np.zeros((n, kmax), np.float64)
    cnt = np.full(n, -1, np.int32)          # -1: no entry in the dict at all
    for i, v in sim.items():
        c = len(v["indexes"])
        idx[i, :c] = v["indexes"]
        val[i, :c] = v["values"]
        cnt[i] = c
    return idx, val, cnt


def inverted_to_csr(inv, n_sets):
    rows = [[] for _ in range(n_sets)]
    for e in sorted(inv):
        for s in inv[e]:
            rows[s].append(e)
    rowptr = np.zeros(n_sets + 1, np.int64)
    for s in range(n_sets):
        rowptr[s + 1] = rowptr[s] + len(rows[s])
    colidx = np.array([e for r_ in rows for e in sorted(r_)], np.int32)
    return rowptr, colidx


# --------------------------------------------------------------------------- C1
def make_c1():
    work = tempfile.mkdtemp(prefix="qrs_golden_")
    shutil.copytree(os.path.join(REF, "resources", "input"), os.path.join(work, "input"))
    os.makedirs(os.path.join(work, "output"))
    cwd = os.getcwd()
    os.chdir(work)
    try:
        random.seed(1234)               # generator.py:39 seeds from the clock at import
        np.random.seed(1234)
        with quiet():
            ref_gen.get_data()
            ref_gen.create_dataset()
            ref_gen.create_users()
            ref_gen.create_queries()
            ref_gen.create_matrix()
        out = os.path.join(HERE, "c1")
        os.makedirs(out, exist_ok=True)
        for f_ in ("dataset", "users", "queries", "utility_matrix"):
            shutil.copy(os.path.join(work, "output", f_ + ".csv"), os.path.join(out, f_ + ".csv"))

        # main.py:12-85, verbatim sequence of calls
        rec = ref_rec.Recommender()
        dataset = dt.fread("./output/dataset.csv")
        rec.datasetFeatures = list(dataset.names)[1::]
        users = dt.fread("./output/users.csv", header=False)
        queries, queries_ids = rec.parse_queries("./output/queries.csv")
        ratings = dt.fread("./output/utility_matrix.csv", columns=["user"] + queries_ids)
        rec.init(users, queries, queries_ids, dataset, ratings)
    finally:
        os.chdir(cwd)

    cap = {}
    orig_shingles = rec.compute_shingles
    orig_sigs = rec.compute_signatures
    orig_q = rec.compute_querySimilarities
    orig_u = rec.compute_userSimilarities
    orig_cand = ref_lsh.LSH.get_candidates

    def shingles():
        cap["inv"] = orig_shingles()
        return cap["inv"]

    def sigs():
        np.random.seed(7)
        cap["sig"] = orig_sigs()
        return cap["sig"]

    def qsim():
        cap["qsim"] = orig_q()
        return cap["qsim"]

    def usim():
        cap["usim"] = orig_u()
        return cap["usim"]

    def cand(self, signatures):
        cap["b"] = self.b
        cap["cand"] = orig_cand(self, signatures)
        return cap["cand"]

    rec.compute_shingles, rec.compute_signatures = shingles, sigs
    rec.compute_querySimilarities, rec.compute_userSimilarities = qsim, usim
    ref_lsh.LSH.get_candidates = cand
    ref_lsh.LSH.buckets = []
    try:
        with quiet():
            to_predict, predictions, missed = rec.compute_scores()
    finally:
        ref_lsh.LSH.get_candidates = orig_cand

    n_sets = rec.queriesIDs.size
    m = rec.dataset.shape[0]
    K = ref_rec.PERM
    b = cap["b"]
    rowptr, colidx = inverted_to_csr(cap["inv"], n_sets)
    bk = buckets_to_arrays(ref_lsh.LSH.buckets, K // b)
    q_idx, q_val, q_cnt = sim_to_arrays(cap["qsim"], n_sets)
    u_idx, u_val, u_cnt = sim_to_arrays(cap["usim"], rec.usersIDs.size)
    cand_arr = np.array(sorted(cap["cand"]), np.int64).reshape(-1, 2)
    np.random.seed(7)
    perm_tab = np.stack([np.random.permutation(m) for _ in range(K)], axis=1).astype(np.int32)
    np.savez_compressed(
        os.path.join(HERE, "c1.npz"),
        n_sets=n_sets, m=m, K=K, b=b, seed=7,
        rowptr=rowptr, colidx=colidx, perm_tab=perm_tab.astype(np.int16),
        sig=np.asarray(cap["sig"]).astype(np.int16),
        bucket_band=bk[0], bucket_lanes=bk[1], bucket_offs=bk[2], bucket_members=bk[3],
        candidates=cand_arr.astype(np.int32),
        q_idx=q_idx, q_val=q_val, q_cnt=q_cnt, u_idx=u_idx, u_val=u_val, u_cnt=u_cnt,
        ratings=rec.ratings.astype(np.int16),
        to_predict=to_predict.astype(np.int32),
        predictions=predictions.to_numpy().astype(np.int16),
        missed=missed.astype(np.int32),
        users_ids=np.array(rec.usersIDs, dtype=str), queries_ids=np.array(rec.queriesIDs, dtype=str),
    )
    shutil.rmtree(work)
    print("c1: n_sets=%d m=%d K=%d b=%d nnz=%d cand=%d to_predict=%d missed=%d"
          % (n_sets, m, K, b, colidx.size, len(cap["cand"]), len(to_predict), len(missed)))


# ------------------------------------------------------------------------ synth
def planted_sets(n, m, density, group, seed, empty_rows=()):
    """Small planted-taste-group generator (SURVEY.md §8d recipe, fixture-local)."""
    rng = np.random.default_rng(seed)
    base = max(1, round(density * m))
    rows = []
    for g0 in range(0, n, group):
        proto = rng.choice(m, size=base, replace=False)
        for _ in range(min(group, n - g0)):
            keep = proto[rng.random(base) < 0.8]
            extra = rng.integers(0, m, size=rng.binomial(base, 0.2))
            rows.append(np.unique(np.concatenate([keep, extra])).astype(np.int32))
    for e in empty_rows:
        rows[e] = np.zeros(0, np.int32)
    rowptr = np.zeros(n + 1, np.int64)
    rowptr[1:] = np.cumsum([r.size for r in rows])
    return rowptr, np.concatenate(rows).astype(np.int32)


def run_reference_lsh(name, n, m, density, group, K, bands, data_seed, perm_seed, empty_rows=(), k=None):
    rowptr, colidx = planted_sets(n, m, density, group, data_seed, empty_rows)
    inv = {e: [] for e in range(m)}
    for s in range(n):
        for e in colidx[rowptr[s]:rowptr[s + 1]]:
            inv[int(e)].append(s)

    class _Shape:
        shape = (m, 1)

    rec = ref_rec.Recommender()
    rec.compute_shingles = lambda: inv
    rec.dataset = _Shape()
    rec.queriesIDs = np.arange(n)
    rec.queries = np.zeros((n, 5))
    ref_rec.PERM = K
    np.random.seed(perm_seed)
    with quiet():
        sig = rec.compute_signatures()
    ref_rec.PERM = 180
    out = dict(n=n, m=m, K=K, perm_seed=perm_seed, rowptr=rowptr, colidx=colidx,
               sig=np.asarray(sig).astype(np.int32), bands=np.array(bands, np.int32))
    kk = k or round(np.log(n) / np.log(1.5))
    out["k"] = kk
    for b in bands:
        ref_lsh.LSH.buckets = []
        l = ref_lsh.LSH(b)
        for s in sig:
            l.compute_buckets(s)
        cand = l.get_candidates(sig)
        bk = buckets_to_arrays(ref_lsh.LSH.buckets, K // b)
        out["b%d_bucket_band" % b], out["b%d_bucket_lanes" % b] = bk[0], bk[1]
        out["b%d_bucket_offs" % b], out["b%d_bucket_members" % b] = bk[2], bk[3]
        cand_arr = np.array(sorted(cand), np.int64).reshape(-1, 2)
        out["b%d_candidates" % b] = cand_arr.astype(np.int32)
        # cosine/top-k block, recommender.py:183-210, driven exactly as the reference does
        from sklearn.metrics.pairwise import cosine_similarity
        qs = {}
        for i, j in cand:
            for a in (i, j):
                if a not in qs:
                    qs[a] = {"indexes": [], "signatures": [sig[a]]}
            qs[i]["indexes"].append(j)
            qs[j]["indexes"].append(i)
            qs[i]["signatures"].append(sig[j])
            qs[j]["signatures"].append(sig[i])
        all_vals = {}
        for i in qs:
            vals = np.around(cosine_similarity(qs[i]["signatures"])[0][1::], 3)
            for j, v in zip(qs[i]["indexes"], vals):
                all_vals[(min(i, j), max(i, j))] = v
            ind = np.argsort(vals)[::-1][0:kk]
            qs[i]["indexes"] = np.array(qs[i]["indexes"])[ind]
            qs[i]["values"] = np.array(vals)[ind]
            del qs[i]["signatures"]
        out["b%d_pair_sims" % b] = np.array([all_vals[tuple(p)] for p in cand_arr.tolist()], np.float64)
        t = sim_to_arrays(qs, n)
        out["b%d_q_idx" % b], out["b%d_q_val" % b], out["b%d_q_cnt" % b] = t
        print("%s: K=%d b=%d buckets=%d cand=%d" % (name, K, b, bk[0].size, len(cand)))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)


def make_wavg():
    rng = np.random.default_rng(99)
    cases = []
    wa = ref_rec.weighted_average
    for t in range(400):
        Q = 40
        vec = rng.integers(0, 101, size=Q).astype(np.int32)
        vec[rng.random(Q) < 0.5] = 0
        if t % 7 == 0:
            vec[vec != 0] = 5          # exact .5 tie of the hybrid rule: 5*0.7+18 = 21.5
        if t % 11 == 0:
            vec[:] = 0
        c = int(rng.integers(1, 12))
        idx = rng.choice(Q, size=c, replace=False).astype(np.int64)
        sims = np.sort(np.around(rng.random(c), 3))[::-1].copy()
        cases.append((vec, idx, sims, float(wa(vec, idx, sims))))
    kmax = 12
    vecs = np.stack([c[0] for c in cases])
    idxs = np.full((len(cases), kmax), -1, np.int32)
    sims = np.zeros((len(cases), kmax), np.float64)
    cnts = np.zeros(len(cases), np.int32)
    for t, c in enumerate(cases):
        idxs[t, :c[1].size] = c[1]
        sims[t, :c[1].size] = c[2]
        cnts[t] = c[1].size
    res = np.array([c[3] for c in cases], np.float64)
    # hybrid rule evaluated by the reference's own expressions (recommender.py:324-331)
    QW, UW, DM = ref_rec.QUERY_WEIGHT, ref_rec.USER_WEIGHT, ref_rec.DEFAULT_MEAN
    hy = []
    for t in range(len(cases)):
        q = res[t]
        u = res[(t * 7 + 3) % len(cases)]
        if u == 0 and q == 0:
            v = 0
        elif u == 0:
            v = round(q * (QW + (UW * 0.5)) + DM * (UW * 0.5))
        elif q == 0:
            v = round(u * (UW + (QW * 0.5)) + DM * (QW * 0.5))
        else:
            v = round(q * QW + u * UW)
        hy.append((q, u, v))
    np.savez_compressed(os.path.join(HERE, "wavg.npz"), vecs=vecs, idxs=idxs, sims=sims, cnts=cnts,
                        result=res, hybrid=np.array(hy, np.float64))
    print("wavg: %d cases" % len(cases))


if __name__ == "__main__":
    make_c1()
    run_reference_lsh("synth_a", n=256, m=600, density=0.05, group=16, K=128, bands=[32, 16, 64],
                      data_seed=20260101, perm_seed=7, empty_rows=(0, 17, 255))
    run_reference_lsh("synth_b", n=192, m=2000, density=0.02, group=8, K=64, bands=[8, 16, 32],
                      data_seed=20260102, perm_seed=11, empty_rows=(5,))
    run_reference_lsh("synth_wrap", n=96, m=70000, density=0.00003, group=8, K=16, bands=[4, 8, 16],
                      data_seed=20260103, perm_seed=13, empty_rows=(3,))
    make_wavg()
