"""TEST DOUBLE for `query_recommendation_system_b200.stages` on CPU tensors.

Test infrastructure only: it lets the multi-rank HOST logic of `pipeline.LshCfPipeline` (row
sharding, signature all-gather, bucket-record all-to-all keyed by the hash bits, pair routing,
ratings all-gather) run under `gloo` with world_size 2 in a container without a GPU.  Every stage
is answered by the CPU oracle (oracle/spec.py) in the SAME data formats the CUDA stages use
(bucket records hash24|band|id, packed pairs, dense neighbour tables, tiled sparse predictions),
so what is tested is the plumbing, not the kernels (those are tested on the GPU against the same
oracle).  Nothing under query_recommendation_system_b200/ imports this module.
"""
import math

import numpy as np
import torch

from oracle import spec

U32 = np.uint32


def band_records_np(sig, b, id_base):
    """common.cuh band_record(): hash24 << 40 | band << 32 | id, sentinel 0xFFFFFF for all-(-1) bands."""
    n, K = sig.shape
    r = K // b
    lanes = (sig.astype(np.int64) & 0xFFFF).astype(U32).reshape(n, b, r)
    with np.errstate(over="ignore"):
        h = (U32(0x9E3779B9) * (np.arange(b, dtype=U32) + U32(1)))[None, :].repeat(n, 0)
        for c in range(0, r, 2):
            w = lanes[:, :, c].copy()
            if c + 1 < r:
                w |= lanes[:, :, c + 1] << U32(16)
            h = ((h ^ w) * U32(0x9E3779B1)).astype(U32)
            h ^= h >> U32(15)
        h = (h * U32(0x85EBCA6B)).astype(U32)
        h ^= h >> U32(16)
        h = h >> U32(8)
    h[h == 0xFFFFFF] = 0xFFFFFE
    h[(lanes == 0xFFFF).all(axis=2)] = 0xFFFFFF
    ids = (np.arange(n, dtype=np.uint64) + np.uint64(id_base))[:, None]
    rec = (h.astype(np.uint64) << np.uint64(40)) | (np.arange(b, dtype=np.uint64)[None, :] << np.uint64(32)) | ids
    return np.ascontiguousarray(rec.T)            # [b, n]


def choose_head_ranks(m, avg_len):
    return 0


def perm_head(perm_tab, R):
    return None


def minhash(rowptr, colidx, m, perm_tab, order_head, b=0, id_base=0, want_norm2=True, mode=0, **kw):
    sig = spec.signatures(rowptr.numpy(), colidx.numpy(), perm_tab.numpy())
    norm2 = torch.from_numpy((sig * sig).sum(axis=1)) if want_norm2 else None
    rec = torch.from_numpy(band_records_np(sig, b, id_base).view(np.int64)) if b > 0 else None
    return torch.from_numpy(sig.astype(np.int32)), norm2, rec


def sort_u64(keys, n_per_seg, segs, bit_lo, bit_hi, tmp=None):
    k = keys.numpy().view(np.uint64).reshape(segs, n_per_seg)
    width = bit_hi - bit_lo
    mask = np.uint64((1 << width) - 1) if width < 64 else np.uint64(2 ** 64 - 1)
    for s in range(segs):
        field = (k[s] >> np.uint64(bit_lo)) & mask
        k[s] = k[s][np.argsort(field, kind="stable")]
    return keys


def pair_slots(records):
    rec = records.numpy().view(np.uint64)
    key = rec >> np.uint64(32)
    n = rec.size
    cnt = np.zeros(n, np.int64)
    for p in range(1, n):
        if key[p] == key[p - 1] and (int(key[p]) >> 8) != 0xFFFFFF:
            cnt[p] = cnt[p - 1] + 1
    off = np.zeros(n + 1, np.int64)
    off[1:] = np.cumsum(cnt)
    return torch.from_numpy(off)


def pairs_score(records, slot_off, sig, norm2, b, capacity, sig_id_base=0, want_sims=True, fill_sentinel=False):
    rec = records.numpy().view(np.uint64)
    sg = sig.numpy().astype(np.int64)
    K = sg.shape[1]
    lanes16 = (sg & 0xFFFF).reshape(sg.shape[0], b, K // b)
    off = slot_off.numpy()
    out = []
    for p in range(rec.size):
        t = int(off[p + 1] - off[p])
        for q in range(p - t, p):
            band = int((rec[p] >> np.uint64(32)) & np.uint64(0xFF))
            i, j = sorted((int(rec[q] & np.uint64(0xFFFFFFFF)), int(rec[p] & np.uint64(0xFFFFFFFF))))
            li, lj = lanes16[i - sig_id_base], lanes16[j - sig_id_base]
            match = (li == lj).all(axis=1) & ~(li == 0xFFFF).all(axis=1)
            if match[band] and not match[:band].any():
                out.append((i, j))
    pairs = np.array(out, np.int64).reshape(-1, 2)
    packed = (pairs[:, 0].astype(np.uint64) << np.uint64(32)) | pairs[:, 1].astype(np.uint64)
    sims = spec.cosine_milli(sg, pairs - sig_id_base) if (want_sims and pairs.size) else np.zeros(pairs.shape[0], np.int32)
    pt = torch.full((max(capacity, 1),), -1 if fill_sentinel else 0, dtype=torch.int64)
    stt = torch.zeros(max(capacity, 1), dtype=torch.int32)
    w = min(packed.size, capacity)                   # like the kernel: writes beyond the capacity are dropped
    pt[:w] = torch.from_numpy(packed.view(np.int64))[:w]
    stt[:w] = torch.from_numpy(sims.astype(np.int32))[:w]
    return pt, stt, torch.tensor([packed.size], dtype=torch.int64)


def center_rows(rowptr, val):
    cval, n2 = spec.centered_rows(rowptr.numpy(), val.numpy())
    return torch.from_numpy(cval), torch.from_numpy(n2)


def pair_centered_cosine(rowptr, colidx, val, cval, norm2, pairs, m, grouped=False, force_search=False):
    p = pairs.numpy().view(np.uint64)
    if grouped:
        assert (np.diff((p >> np.uint64(32)).astype(np.int64)) >= 0).all(), "pairs are not grouped by first id"
    pr = np.stack([(p >> np.uint64(32)).astype(np.int64), (p & np.uint64(0xFFFFFFFF)).astype(np.int64)], 1)
    return torch.from_numpy(spec.centered_cosine_milli(rowptr.numpy(), colidx.numpy(), val.numpy(), pr))


def topk_tables(pairs, sims, n_pairs, row_base, n_rows, k, want_milli=False):
    p = pairs.numpy().view(np.uint64)[:n_pairs]
    keep = p != np.uint64(2 ** 64 - 1)               # sentinel entries (empty exchange / buffer slots) are skipped
    sims = sims[:n_pairs][torch.from_numpy(keep)]
    p = p[keep]
    n_pairs = p.size
    pr = np.stack([(p >> np.uint64(32)).astype(np.int64), (p & np.uint64(0xFFFFFFFF)).astype(np.int64)], 1)
    # the same pair can arrive once per owned end; the kernel counts each (row, neighbour) edge once per copy,
    # so the router must not duplicate — assert it here
    assert np.unique(p).size == p.size, "a routed pair arrived twice on one rank"
    n_all = max(row_base + n_rows, int(pr.max()) + 1 if pr.size else 0)
    idx, val, cnt = spec.topk_tables(n_all, pr, sims.numpy()[:n_pairs] / 1000.0, k)
    sl = slice(row_base, row_base + n_rows)
    out = (torch.from_numpy(idx[sl].copy()), torch.from_numpy(val[sl].copy()), torch.from_numpy(cnt[sl].copy()))
    if want_milli:
        out = out + (torch.from_numpy(np.rint(val[sl] * 1000).astype(np.int32)),)
    return out


def choose_predict_tile(Q, avg_len=0.0, integer_tables=True):
    return max(32, int(math.ceil(Q / 3 / 32.0)) * 32)


def predict_workspace_bytes(U, nnz, Q, tile, capacity):
    return 0


class PeerWorkspace:
    """CPU stand-in: nothing is mapped; the plan below gathers the shards with gloo instead."""

    def __init__(self, nbytes, group=None):
        import torch.distributed as dist
        self.nbytes, self.group = nbytes, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)

    def close(self):
        pass


class PredictPlan:
    """Nothing to prepare on the CPU double.  With peers the whole matrix is gathered (test infrastructure: the
    product reads remote rows in place through CUDA IPC mappings), so neighbour ids are GLOBAL row ids."""

    def __init__(self, rowptr, colidx, val, Q, tile, capacity_hint=0, peers=None):
        self.tile, self.Q, self.capacity, self.peers = tile, Q, capacity_hint, peers
        self.row_base = 0
        self.csr = (rowptr, colidx, val)
        if peers is not None:
            import torch.distributed as dist
            parts = [None] * peers.world
            dist.all_gather_object(parts, (rowptr.numpy(), colidx.numpy(), val.numpy()), group=peers.group)
            n_local = rowptr.numel() - 1
            lens = np.concatenate([np.diff(p[0]) for p in parts])
            rp = np.zeros(lens.size + 1, np.int64)
            np.cumsum(lens, out=rp[1:])
            self.csr = (torch.from_numpy(rp), torch.from_numpy(np.concatenate([p[1] for p in parts])),
                        torch.from_numpy(np.concatenate([p[2] for p in parts])))
            self.row_base = peers.rank * n_local


def predict_sparse(rowptr, colidx, val, Q, u_tab, user_base, n_users, tile, capacity, u_milli=None, n_regions=1,
                   plan=None, **kw):
    assert n_regions == 1
    if plan is not None and plan.peers is not None:
        rowptr, colidx, val = plan.csr
        user_base = user_base + plan.row_base
    idx, v, cnt = [x.numpy() for x in u_tab]
    full_idx = np.full((user_base + n_users, idx.shape[1]), -1, np.int32)
    full_val = np.zeros((user_base + n_users, idx.shape[1]))
    full_cnt = np.zeros(user_base + n_users, np.int32)
    full_idx[user_base:], full_val[user_base:], full_cnt[user_base:] = idx, v, cnt
    ref = spec.predict_user_branch_csr(rowptr.numpy(), colidx.numpy(), val.numpy(), Q, full_idx, full_val, full_cnt,
                                       rows=range(user_base, user_base + n_users))
    nt = (Q + tile - 1) // tile
    rs = np.zeros((n_users, nt), np.int64)
    rc = np.zeros((n_users, nt), np.int32)
    cols, preds = [], []
    pos = 0
    for t in range(n_users):
        c, p = ref[user_base + t]
        for ti in range(nt):
            sel = (c >= ti * tile) & (c < (ti + 1) * tile)
            rs[t, ti], rc[t, ti] = pos, int(sel.sum())
            cols.append(c[sel])
            preds.append(p[sel])
            pos += int(sel.sum())
    cols = np.concatenate(cols) if cols else np.zeros(0, np.int32)
    preds = np.concatenate(preds) if preds else np.zeros(0, np.int64)
    ct = torch.zeros(max(capacity, 1), dtype=torch.int32)
    pt = torch.zeros(max(capacity, 1), dtype=torch.uint8)
    w = min(capacity, pos)
    ct[:w] = torch.from_numpy(cols[:w].astype(np.int32))
    pt[:w] = torch.from_numpy(preds[:w].astype(np.uint8))
    return (torch.from_numpy(rs), torch.from_numpy(rc), ct, pt, torch.tensor([pos], dtype=torch.int64),
            torch.tensor([pos], dtype=torch.int64))


def group_pad(sorted_records, lg, cap):
    rec = sorted_records.numpy().view(np.uint64)
    G = 1 << lg
    dest = (rec >> np.uint64(64 - lg)).astype(np.int64)
    assert (np.diff(dest) >= 0).all(), "records are not grouped by destination"
    out = np.full(G * cap, 2 ** 64 - 1, np.uint64)
    counts = np.bincount(dest, minlength=G).astype(np.int64)
    start = np.concatenate([[0], np.cumsum(counts)])
    for d in range(G):
        w = min(cap, int(counts[d]))
        out[d * cap:d * cap + w] = rec[start[d]:start[d] + w]
    over = np.array([1 if (counts > cap).any() else 0], np.int32)
    return torch.from_numpy(out.view(np.int64)), torch.from_numpy(counts), torch.from_numpy(over)


def route_pairs(pairs, sims, counter, rows_per_rank, n_ranks, cap):
    n = min(int(counter.item()), pairs.numel())
    p = pairs.numpy().view(np.uint64)[:n]
    i, j = (p >> np.uint64(32)).astype(np.int64), (p & np.uint64(0xFFFFFFFF)).astype(np.int64)
    word = (i.astype(np.uint64) << np.uint64(37)) | (j.astype(np.uint64) << np.uint64(11)) | \
        (sims.numpy()[:n].astype(np.int64) + 1000).astype(np.uint64)
    oi, oj = i // rows_per_rank, j // rows_per_rank
    out = np.full(n_ranks * cap, 2 ** 64 - 1, np.uint64)
    counts = np.zeros(n_ranks, np.int64)
    over = 0
    for d in range(n_ranks):
        sel = (oi == d) | (oj == d)
        w = word[sel]
        counts[d] = w.size
        if w.size > cap:
            over = 2
        out[d * cap:d * cap + min(cap, w.size)] = w[:cap]
    return torch.from_numpy(out.view(np.int64)), torch.from_numpy(counts), torch.tensor([over], dtype=torch.int32)


def unpack_routed(words):
    w = words.numpy().view(np.uint64)
    sent = w == np.uint64(2 ** 64 - 1)
    pairs = ((w >> np.uint64(37)) << np.uint64(32)) | ((w >> np.uint64(11)) & np.uint64(0x3FFFFFF))
    pairs[sent] = np.uint64(2 ** 64 - 1)
    sims = (w & np.uint64(0x7FF)).astype(np.int64) - 1000
    sims[sent] = 0
    return torch.from_numpy(pairs.view(np.int64)), torch.from_numpy(sims.astype(np.int32))
