"""Drop-in for the reference's `recommender.Recommender` (recommender.py:49-416) with the hot
path on the B200 kernels.

Surface kept (what main.py:12-114 uses): `Recommender()`, attribute `datasetFeatures`,
`parse_queries(path)`, `init(users, queries, queriesIDs, dataset, ratings)`,
`compute_scores() -> (to_predict, predictions DataFrame, missed)`, `top_k_queries(...)`, plus the
internal stage methods `compute_shingles / compute_signatures / compute_querySimilarities /
compute_userSimilarities` with their reference return types, and the module constants.

On the GPU: signatures (a1), banding + grouping + candidate pairs + cosine (a2-a8), top-k neighbour
tables (a9), the per-cell weighted averages + hybrid rule (a10-a11) and — the first "next" row, f-1 —
the predicate evaluation of compute_shingles on dictionary-encoded columns; with user_branch="lsh" also the
user-similarity tables (f-3: LSH candidates over the rating rows + the reference's centred cosine).  On the host, as in the
reference and out of this path's scope: the PCA + BIRCH user clustering
(compute_userSimilarities), CSV parsing and the interactive prompt.  The reference reads its tables through `datatable`, which is
not installable here; `init` accepts anything with `.to_numpy()` (datatable Frames, pandas
DataFrames) or plain arrays.
"""
from __future__ import annotations

import math
import time

import numpy as np
import pandas as pd
import torch

from . import stages
from ._lib import QrsError
from .lsh import LSH, _device

# constants (recommender.py:30-34, 153)
PERM = 180            # number of independent hash functions
QUERY_WEIGHT = 0.6
USER_WEIGHT = 0.4
DEFAULT_MEAN = 60
LSH_THRESH = 0.2


def weighted_average(ratings, indexes, simScores):
    """recommender.py:36-47 for callers that use the module function directly (scalar helper; the
    kernels compute the same expression for every cell at once in qrs_predict_*)."""
    num = 0.0
    wsum = 0.0
    for t in range(len(indexes)):
        r = int(ratings[int(indexes[t])])
        s = float(simScores[t])
        num = num + r * s
        if r != 0:
            wsum = wsum + s
    return num / wsum if wsum != 0 else 0


def choose_bands(K: int, thresh: float = LSH_THRESH):
    """recommender.py:156-163: the largest b | K with b % 10 == 0 and round((1/b)**(1/r), 2) >= thresh;
    None where the reference would fail with an unbound `band` (e.g. K = 128)."""
    for b in range(K, 0, -1):
        if K % b == 0 and b % 10 == 0:
            r = K / b
            if round((1 / b) ** (1 / r), 2) >= thresh:
                return b
    return None


def _np(obj):
    return obj.to_numpy() if hasattr(obj, "to_numpy") else np.asarray(obj)


def _table_to_dict(idx, val, cnt, only_nonempty: bool):
    """Dense neighbour table → the reference's {i: {"indexes": int64[], "values": float64[]}}."""
    out = {}
    idx, val, cnt = idx.cpu().numpy(), val.cpu().numpy(), cnt.cpu().numpy()
    for i in range(idx.shape[0]):
        c = int(cnt[i])
        if c == 0 and only_nonempty:
            continue                                  # recommender.py:186-196: no entry without a candidate
        out[i] = {"indexes": idx[i, :c].astype(np.int64), "values": val[i, :c].copy()}
    return out


def _dict_to_table(sim: dict, n: int, device):
    k = max([len(v["indexes"]) for v in sim.values()] + [1])
    idx = np.full((n, k), -1, np.int32)
    val = np.zeros((n, k), np.float64)
    cnt = np.zeros(n, np.int32)
    for i, v in sim.items():
        c = len(v["indexes"])
        idx[i, :c] = v["indexes"]
        val[i, :c] = v["values"]
        cnt[i] = c
    return tuple(torch.from_numpy(a).to(device) for a in (idx, val, cnt))


class Recommender:
    def __init__(self, device=None, bands=None, max_candidates=None, verbose: bool = True,
                 user_branch: str = "cluster"):
        """`bands` / `max_candidates` override the reference's rules (needed when PERM has no
        divisor that is a multiple of 10, recommender.py:156-165).  `user_branch`: "cluster" = the
        reference's PCA + BIRCH pipeline on the host; "lsh" = LSH candidates over the users' rating rows
        scored on the GPU with the same centred cosine (for shapes the dense pipeline
        cannot reach)."""
        if user_branch not in ("cluster", "lsh"):
            raise ValueError("user_branch must be 'cluster' or 'lsh'")
        self._user_branch = user_branch
        self.datasetFeatures = []
        self.tupleCount = {}
        self._device_arg = device
        self._bands = bands
        self._kmax = max_candidates
        self._verbose = verbose
        self._sig_dev = None

    def _log(self, *a):
        if self._verbose:
            print(*a)

    # ------------------------------------------------------------------ I/O glue (recommender.py:51-64)
    def init(self, users, queries, queriesIDs, dataset, ratings):
        self.usersIDs = _np(users).T[0]
        self.queries = _np(queries)
        self.queriesIDs = np.array(queriesIDs)
        ds = dataset.to_pandas() if hasattr(dataset, "to_pandas") else pd.DataFrame(dataset)
        self.dataset = ds.astype(str)
        self.tupleCount = {}
        if isinstance(ratings, pd.DataFrame):
            r = ratings.drop(columns=["user"]) if "user" in ratings.columns else ratings
            r = r.to_numpy(dtype=np.float64, na_value=0.0) if hasattr(r, "to_numpy") else np.asarray(r)
        elif hasattr(ratings, "to_numpy"):
            names = list(getattr(ratings, "names", []))
            r = ratings.to_numpy()
            if "user" in names:
                r = np.delete(r, names.index("user"), axis=1)
        else:
            r = np.asarray(ratings)
        r = np.where(pd.isna(r), 0, r) if r.dtype.kind in "fO" else r
        self.ratings = np.asarray(r).astype(np.int32)

    def parse_queries(self, path: str):
        """recommender.py:386-416: `Q12,attr=val,...` lines → (frame aligned to datasetFeatures, ids)."""
        data, indexes = [], []
        with open(path) as f:
            for row in f:
                values = row.rstrip("\n").split(",")
                indexes.append(values[0])
                element = ["" for _ in self.datasetFeatures]
                for cond in values[1:]:
                    name, value = cond.split("=")
                    element[self.datasetFeatures.index(name)] = value
                data.append(element)
        frame = pd.DataFrame(data, columns=list(self.datasetFeatures), dtype=str) if data else pd.DataFrame()
        return frame, indexes

    # ------------------------------------------------------------------ set construction (f-1)
    def compute_shingles(self):
        """recommender.py:68-103: answer set of every conjunctive query → {tuple index: [query ids]}.
        The columns are dictionary-encoded on the host once; the predicate evaluation (the reference's pandas
        boolean masks, one per query and attribute) runs on the GPU (qrs_answer_sets_*) and leaves the
        sets × elements CSR (query × tuple) the signature kernel consumes."""
        drows = self.dataset.shape[0]
        nq = self.queriesIDs.size
        F = len(self.datasetFeatures)
        dev = _device(self._device_arg)
        table = np.empty((F, drows), dtype=np.int32)
        qcodes = np.full((nq, F), -1, dtype=np.int32)
        for fi, ft in enumerate(self.datasetFeatures):
            values, inverse = np.unique(self.dataset[ft].to_numpy().astype(str), return_inverse=True)
            table[fi] = inverse
            code_of = {v: c for c, v in enumerate(values.tolist())}
            for q in range(nq):
                want = self.queries[q][fi]
                if want != "":
                    qcodes[q, fi] = code_of.get(str(want), -2)          # -2: the value never occurs → empty answer
        rowptr_d, colidx_d = stages.answer_sets(torch.from_numpy(table).to(dev), torch.from_numpy(qcodes).to(dev))
        self._sets_dev = (rowptr_d, colidx_d)
        self._sets_rowptr = rowptr_d.cpu().numpy()
        self._sets_colidx = colidx_d.cpu().numpy()
        shingles = {d: [] for d in range(drows)}
        for q in range(nq):
            answer = self._sets_colidx[self._sets_rowptr[q]:self._sets_rowptr[q + 1]]
            self.tupleCount[q] = answer.size
            for ind in answer:
                shingles[int(ind)].append(q)
        return shingles

    def set_sets_csr(self, rowptr, colidx, n_elements: int):
        """Binding seam (SURVEY.md §8 preamble): feed sets × elements directly (e.g. users × rated
        queries) instead of evaluating string predicates."""
        self._sets_rowptr = np.asarray(rowptr, np.int64)
        self._sets_colidx = np.asarray(colidx, np.int32)
        self._sets_dev = None
        self._n_elements = int(n_elements)

    # ------------------------------------------------------------------ a1
    def compute_signatures(self):
        """recommender.py:105-143 → int64[n_sets, PERM]; PERM successive np.random.permutation draws
        from the global legacy RNG exactly like the reference (seed it to reproduce a run)."""
        if getattr(self, "_n_elements", None) is not None:
            m = self._n_elements                       # sets injected through set_sets_csr
        else:
            self.compute_shingles()
            m = self.dataset.shape[0]
        K = PERM
        dev = _device(self._device_arg)
        t0 = time.time()
        tab = np.empty((m, K), dtype=np.int32)
        for i in range(K):
            tab[:, i] = np.random.permutation(m)                        # recommender.py:120
        if getattr(self, "_sets_dev", None) is not None:
            rowptr, colidx = self._sets_dev                              # left on the device by compute_shingles
        else:
            rowptr = torch.from_numpy(self._sets_rowptr).to(dev)
            colidx = torch.from_numpy(self._sets_colidx).to(dev)
        n = rowptr.numel() - 1
        tab_d = torch.from_numpy(tab).to(dev)
        R = stages.choose_head_ranks(m, colidx.numel() / max(n, 1))
        head = stages.perm_head(tab_d, R) if R > 0 else None
        sig, norm2, _ = stages.minhash(rowptr, colidx, m, tab_d, head, b=0, want_norm2=True)
        self._sig_dev, self._norm2_dev = sig, norm2
        out = sig.cpu().numpy().astype(np.int64)
        self._log(str(round(time.time() - t0, 3)) + "s for signature_matrix")
        return out

    # ------------------------------------------------------------------ a2-a9
    def compute_querySimilarities(self):
        """recommender.py:145-214 → {i: {"indexes", "values"}} for sets with >= 1 candidate, neighbours in
        canonical order (similarity desc, id asc; the reference's tie order is unspecified)."""
        t0 = time.time()
        self.compute_signatures()
        sig = self._sig_dev
        n, K = sig.shape
        kmax = self._kmax if self._kmax is not None else round(math.log(n, 1.5))        # recommender.py:151
        band = self._bands if self._bands is not None else choose_bands(K)
        if band is None:
            raise QrsError("no band count satisfies the reference rule for PERM=%d (recommender.py:156-165); "
                           "pass bands= explicitly" % K)
        if kmax > 64:
            self._log("MAX_CANDIDATES %d clamped to 64 (the top-k kernel keeps 64 keys per row)" % kmax)
        kmax = max(1, min(kmax, 64))
        lsh = LSH(band, device=sig.device)
        lsh.compute_buckets_batch(sig)                                                    # recommender.py:171-174
        pairs, sims = lsh.candidates_scored(True)                                         # recommender.py:176, 203-204
        self._log("Candidate pairs [{}s]: {}".format(round(time.time() - t0, 3), pairs.numel()))
        idx, val, cnt = stages.topk_tables(pairs, sims, pairs.numel(), 0, n, kmax)        # recommender.py:207-210
        self._q_table = (idx, val, cnt)
        self._candidates = pairs
        self._log(str(round(time.time() - t0, 3)) + "s for overall queries_similarity scores")
        return _table_to_dict(idx, val, cnt, only_nonempty=True)

    # ------------------------------------------------------------------ user branch (host, out of scope)
    def compute_userSimilarities_lsh(self):
        """f-3: the user branch without the dense PCA + BIRCH step.  Users are the sets, their rated queries
        the elements (Binding B); candidates = users sharing an LSH bucket; each candidate pair is scored
        like the reference scores a pair inside a cluster (recommender.py:263-269: cosine of the
        mean-centred, integer-truncated rating rows, 3 decimals, negatives -> 0); top-k as
        recommender.py:274-286 in canonical order.  Every user gets an entry (possibly empty)."""
        t0 = time.time()
        dev = _device(self._device_arg)
        R = np.asarray(self.ratings)
        U, Q = R.shape
        kmax = self._kmax if self._kmax is not None else round(math.log(U, 1.5))          # recommender.py:220
        kmax = max(1, min(kmax, 64))
        band = self._bands if self._bands is not None else choose_bands(PERM)
        if band is None:
            raise QrsError("no band count satisfies the reference rule for PERM=%d; pass bands=" % PERM)
        rows, cols = np.nonzero(R)
        rowptr = np.zeros(U + 1, dtype=np.int64)
        np.cumsum(np.bincount(rows, minlength=U), out=rowptr[1:])
        tab = np.empty((Q, PERM), dtype=np.int32)
        for i in range(PERM):
            tab[:, i] = np.random.permutation(Q)
        rowptr_d = torch.from_numpy(rowptr).to(dev)
        colidx_d = torch.from_numpy(cols.astype(np.int32)).to(dev)
        val_d = torch.from_numpy(R[rows, cols].astype(np.uint8)).to(dev)
        tab_d = torch.from_numpy(tab).to(dev)
        head_ranks = stages.choose_head_ranks(Q, cols.size / max(U, 1))
        head = stages.perm_head(tab_d, head_ranks) if head_ranks > 0 else None
        sig, _, _ = stages.minhash(rowptr_d, colidx_d, Q, tab_d, head, b=0, want_norm2=False)
        lsh = LSH(band, device=sig.device)
        lsh.compute_buckets_batch(sig)
        pairs, _ = lsh.candidates_scored(False)
        P = pairs.numel()
        cval, norm2 = stages.center_rows(rowptr_d, val_d)
        if P:
            stages.sort_u64(pairs, P, 1, 32, 32 + max(1, (U - 1).bit_length()))
        sims = stages.pair_centered_cosine(rowptr_d, colidx_d, val_d, cval, norm2, pairs, Q, grouped=True)
        idx, val, cnt = stages.topk_tables(pairs, sims, P, 0, U, kmax)
        self._u_table = (idx, val, cnt)
        self._log("Candidate user pairs: {}".format(P))
        self._log("\n" + str(round(time.time() - t0, 3)) + "s for overall users_similarity scores")
        return _table_to_dict(idx, val, cnt, only_nonempty=False)

    def compute_userSimilarities(self):
        """recommender.py:216-290 (StandardScaler → PCA(<=200) → BIRCH → per-cluster mean-centred cosine →
        top-k).  Dense sklearn pipeline kept on the host exactly as in the reference: it is not part
        of the LSH path; its output table feeds the GPU prediction kernel.
        With user_branch="lsh" the GPU substitute compute_userSimilarities_lsh runs instead."""
        if self._user_branch == "lsh":
            return self.compute_userSimilarities_lsh()
        from sklearn.cluster import Birch
        from sklearn.decomposition import PCA
        from sklearn.metrics.pairwise import cosine_similarity
        from sklearn.preprocessing import StandardScaler

        n_users = self.usersIDs.size
        kmax = round(math.log(n_users, 1.5))
        n_clusters = round(n_users ** (1 / 1.3))
        scores = StandardScaler().fit_transform(self.ratings)
        comps = min(scores.shape[0], scores.shape[1], 200)
        scores = PCA(n_components=comps).fit(scores).transform(scores)
        labels = Birch(n_clusters=n_clusters).fit(scores).predict(scores)
        sizes = np.bincount(labels)
        for c in range(sizes.size):                      # singletons are pooled into one extra cluster
            if sizes[c] == 1:
                labels[labels == c] = n_clusters
        per_cluster = {}
        for c in np.unique(labels):
            members = np.where(labels == c)[0]
            block = np.array(self.ratings[members])
            for row in block:
                rated = row != 0
                row[rated] = row[rated] - np.mean(row[rated])
            sim = np.around(cosine_similarity(block), 3)
            np.fill_diagonal(sim, 0)
            sim[sim < 0] = 0
            per_cluster[c] = (members, sim)
        out = {}
        for i in range(n_users):
            members, sim = per_cluster[labels[i]]
            me = np.where(members == i)[0][0]
            top = np.argsort(sim[me])[::-1][0:kmax]
            out[i] = {"indexes": np.array(members)[top], "values": np.array(sim[me])[top]}
        return out

    # ------------------------------------------------------------------ a10 + a11
    def compute_scores(self):
        """recommender.py:292-343 → (scores_to_predict int[P, 2], predictions DataFrame[int], missed int[M, 2])."""
        if (QUERY_WEIGHT, USER_WEIGHT, DEFAULT_MEAN) != (0.6, 0.4, 60):
            raise QrsError("the prediction kernel is built with the reference's weights (0.6, 0.4, 60)")
        # limits of the dense drop-in kernel, checked before any work is done (INTEGRATION.md "Limits")
        n_q = len(self.queriesIDs)
        if n_q * 17 > 200 * 1024:
            raise QrsError("compute_scores() keeps one user's whole query axis in shared memory: at most %d queries, "
                           "got %d — use pipeline.LshCfPipeline (sparse predictor) at this scale" % (200 * 1024 // 17, n_q))
        if self.ratings.min() < 0 or self.ratings.max() > 255:
            raise QrsError("ratings must be in 0..255 (generator.py uses 1..100)")
        self._log("\n========== QUERY SIMILARITY ==========")
        self.compute_querySimilarities()
        self._log("\n========== USER SIMILARITY ==========")
        user_sim = self.compute_userSimilarities()
        self._log("\n========== WEIGHTED AVERAGES ==========")
        t0 = time.time()
        final = self.predict_with_tables(self._q_table, user_sim)
        self._log(str(round(time.time() - t0, 3)) + "s for weighted averages")
        scores_to_predict = np.array(np.where(self.ratings == 0)).T
        frame = pd.DataFrame(final, columns=self.queriesIDs, index=self.usersIDs).astype(int)
        missed = np.array(np.where(frame == 0)).T
        return scores_to_predict, frame, missed

    def predict_with_tables(self, query_sim, user_sim):
        """The cell loop of compute_scores for given neighbour tables (dicts in the reference's format or
        dense (idx, val, cnt) device tables; None = branch absent) → int64[users, queries]."""
        dev = _device(self._device_arg)
        R = self.ratings
        if R.min() < 0 or R.max() > 255:
            raise QrsError("ratings must be in 0..255 (generator.py uses 1..100)")
        U, Q = R.shape
        rowptr = np.zeros(U + 1, np.int64)
        rowptr[1:] = np.cumsum((R != 0).sum(axis=1))
        rows, cols = np.nonzero(R)
        to_dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        q_tab = _dict_to_table(query_sim, Q, dev) if isinstance(query_sim, dict) else query_sim
        u_tab = _dict_to_table(user_sim, U, dev) if isinstance(user_sim, dict) else user_sim
        out = stages.predict_dense(to_dev(rowptr), to_dev(cols.astype(np.int32)), to_dev(R[rows, cols].astype(np.uint8)),
                                   Q, q_tab, u_tab)
        return out.cpu().numpy().astype(np.int64)

    # ------------------------------------------------------------------ prompt (recommender.py:345-381)
    def top_k_queries(self, to_predict, predictions, missed):
        predictions = predictions.to_numpy()
        command = ""
        while command.lower() != "no":
            while True:
                user = input("Enter user ID: [int][Max: " + str(self.usersIDs.size) + "] ")
                if user.isdigit() and 0 <= int(user) - 1 < self.usersIDs.size:
                    user = int(user) - 1
                    break
            just_scored = [j for i, j in to_predict if i == user and predictions[i][j] != 0]
            while True:
                k = input("Enter number of recommendations: [int][Max: " + str(len(just_scored)) + "] ")
                if k.isdigit() and 0 < int(k) <= len(just_scored):
                    k = int(k)
                    break
            best = np.argsort(predictions[user][just_scored])[::-1][0:k]
            print("\nTop {} unrated query recommendations for U{}: ".format(k, user + 1))
            for pos in range(best.size):
                print("{}. Q{} - {}".format(pos + 1, just_scored[best[pos]] + 1, predictions[user][just_scored][best[pos]]))
            print()
            command = ""
            while command.lower() not in ["yes", "no"]:
                command = input("Do you want more suggestions? [Yes-No][Default: Yes] ")
                if command == "":
                    command = "yes"
