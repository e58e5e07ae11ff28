"""Stage-level host wrappers: torch CUDA tensors in, torch CUDA tensors out, each a
thin call into the C ABI (include/qrs_b200.h) on torch's current stream.

torch is used for device memory and streams only; every computation below is a
hand-written sm_100a kernel inside libqrs_b200.so.  u64 device data (bucket
records, packed pairs) is carried in int64 tensors (same bits).
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import check, ptr


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _need_cuda(*tensors):
    for t in tensors:
        if t is not None and (not t.is_cuda or not t.is_contiguous()):
            raise _lib.QrsError("stage inputs must be contiguous CUDA tensors (no CPU fallback exists)")


def _ws(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


# --------------------------------------------------------------------------- a1
def choose_head_ranks(m: int, avg_len: float) -> int:
    """Ranks of each permutation kept in the first-touch head table: enough that a
    set of average length misses it with probability < 1e-6, in [32, 2048], <= m."""
    if m < 32:
        return 0
    p = min(max(avg_len, 1.0) / m, 0.999)
    need = math.log(1e-6) / math.log1p(-p)
    r = int(min(max(32, math.ceil(need / 32.0) * 32), 2048))
    return min(r, (m // 32) * 32)


def perm_head(perm_tab: torch.Tensor, R: int) -> Optional[torch.Tensor]:
    """First R ranks of each permutation (recommender.py:121's argsort) in the library's chunk-major layout:
    int32[R / 8, Kp, 8] with Kp = K rounded up to 32 (include/qrs_b200.h, qrs_perm_head_invert)."""
    _need_cuda(perm_tab)
    m, K = perm_tab.shape
    if R <= 0:
        return None
    assert perm_tab.dtype == torch.int32 and R % 8 == 0
    lib = _lib.load()
    Kp = lib.qrs_perm_head_elems(K, R) // R
    out = torch.empty((R // 8, Kp, 8), dtype=torch.int32, device=perm_tab.device)
    check(_lib.load().qrs_perm_head_invert(ptr(perm_tab), m, K, R, ptr(out), _stream()), "qrs_perm_head_invert")
    return out


def minhash(rowptr: torch.Tensor, colidx: torch.Tensor, m: int, perm_tab: torch.Tensor,
            order_head: Optional[torch.Tensor], b: int = 0, id_base: int = 0,
            want_norm2: bool = True, mode: int = _lib.MINHASH_AUTO,
            out_sig: Optional[torch.Tensor] = None, out_norm2: Optional[torch.Tensor] = None,
            out_records: Optional[torch.Tensor] = None):
    """Recommender.compute_signatures (recommender.py:105-143) → sig int32[n, K]; with b > 0
    also the bucket records int64[b, n] of LSH.compute_buckets (lsh.py:31-38)."""
    _need_cuda(rowptr, colidx, perm_tab, order_head)
    assert rowptr.dtype == torch.int64 and colidx.dtype == torch.int32 and perm_tab.dtype == torch.int32
    n = rowptr.numel() - 1
    K = perm_tab.shape[1]
    assert perm_tab.shape[0] == m
    dev = rowptr.device
    sig = out_sig if out_sig is not None else torch.empty((n, K), dtype=torch.int32, device=dev)
    norm2 = None
    if want_norm2:
        norm2 = out_norm2 if out_norm2 is not None else torch.empty(n, dtype=torch.int64, device=dev)
    records = None
    if b > 0:
        records = out_records if out_records is not None else torch.empty((b, n), dtype=torch.int64, device=dev)
    R = 0 if order_head is None else order_head.shape[0] * 8
    check(_lib.load().qrs_minhash_perm(ptr(rowptr), ptr(colidx), n, m, ptr(perm_tab), ptr(order_head), R, K, b,
                                       id_base, ptr(sig), ptr(norm2), ptr(records), n, mode, _stream()),
          "qrs_minhash_perm")
    return sig, norm2, records


def widen_cols_u16(cols16: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """16-bit column ids (the compact wire format of a matrix with at most 65536 queries) → the int32 colidx."""
    _need_cuda(cols16, out)
    assert cols16.dtype in (torch.uint16, torch.int16)
    n = cols16.numel()
    if out is None:
        out = torch.empty(n, dtype=torch.int32, device=cols16.device)
    assert out.dtype == torch.int32 and out.numel() >= n
    check(_lib.load().qrs_cols_widen_u16(ptr(cols16), n, ptr(out), _stream()), "qrs_cols_widen_u16")
    return out[:n]


# ------------------------------------------------------------------------ a2/a3
def band_keys(sig: torch.Tensor, b: int):
    """Exact bucket ids (lsh.py:17-38): keys int64[n, b] (+ keys_hi for r > 4), skipped uint8[n, b]."""
    _need_cuda(sig)
    n, K = sig.shape
    if K % b != 0:
        raise AssertionError("signature length %d is not a multiple of b=%d" % (K, b))   # lsh.py:20
    r = K // b
    keys = torch.empty((n, b), dtype=torch.int64, device=sig.device)
    hi = torch.empty((n, b), dtype=torch.int64, device=sig.device) if r > 4 else None
    skipped = torch.empty((n, b), dtype=torch.uint8, device=sig.device)
    check(_lib.load().qrs_band_keys(ptr(sig), n, K, b, ptr(keys), ptr(hi), ptr(skipped), _stream()), "qrs_band_keys")
    return keys, hi, skipped


def band_records(sig: torch.Tensor, b: int, id_base: int = 0, want_norm2: bool = True):
    _need_cuda(sig)
    n, K = sig.shape
    if K % b != 0:
        raise AssertionError("signature length %d is not a multiple of b=%d" % (K, b))
    records = torch.empty((b, n), dtype=torch.int64, device=sig.device)
    norm2 = torch.empty(n, dtype=torch.int64, device=sig.device) if want_norm2 else None
    check(_lib.load().qrs_band_records(ptr(sig), n, K, b, id_base, ptr(records), n, ptr(norm2), _stream()),
          "qrs_band_records")
    return records, norm2


# ------------------------------------------------------------------------ a3/a4
def sort_u64(keys: torch.Tensor, n_per_seg: int, segs: int, bit_lo: int, bit_hi: int,
             tmp: Optional[torch.Tensor] = None) -> torch.Tensor:
    """In-place stable radix sort of `segs` segments of `n_per_seg` u64 keys on bits [bit_lo, bit_hi)."""
    _need_cuda(keys, tmp)
    assert keys.dtype == torch.int64 and keys.numel() == n_per_seg * segs
    lib = _lib.load()
    if tmp is None:
        tmp = torch.empty_like(keys)
    nbytes = lib.qrs_sort_workspace_bytes(n_per_seg, segs)
    ws = _ws(nbytes, keys.device)
    check(lib.qrs_sort_u64(ptr(keys), ptr(tmp), n_per_seg, segs, bit_lo, bit_hi, ptr(ws), ws.numel(), _stream()),
          "qrs_sort_u64")
    return keys


# --------------------------------------------------------------------------- a5
def pair_slots(records: torch.Tensor) -> torch.Tensor:
    """slot_off int64[n_records + 1]; slot_off[-1] = Σ_runs h(h-1)/2 (lsh.py:49's combinations)."""
    _need_cuda(records)
    lib = _lib.load()
    n = records.numel()
    slot_off = torch.empty(n + 1, dtype=torch.int64, device=records.device)
    ws = _ws(lib.qrs_pair_slots_workspace_bytes(n), records.device)
    check(lib.qrs_pair_slots(ptr(records), n, ptr(slot_off), ptr(ws), ws.numel(), _stream()), "qrs_pair_slots")
    return slot_off


def pairs_score(records: torch.Tensor, slot_off: torch.Tensor, sig: torch.Tensor, norm2: Optional[torch.Tensor],
                b: int, capacity: int, sig_id_base: int = 0, want_sims: bool = True, fill_sentinel: bool = False):
    """LSH.get_candidates (lsh.py:40-54) + rounded cosine (recommender.py:203-204).
    Returns (pairs int64[capacity], sim_milli int32[capacity] | None, counter int64[1] on device).
    fill_sentinel: entries beyond the pairs found are all-ones (a pair no table owns: topk_tables skips them), so
    the whole buffer can be handed on without reading the counter back."""
    _need_cuda(records, slot_off, sig, norm2)
    dev = records.device
    K = sig.shape[1]
    if fill_sentinel:
        pairs = torch.full((max(capacity, 1),), -1, dtype=torch.int64, device=dev)
    else:
        pairs = torch.empty(max(capacity, 1), dtype=torch.int64, device=dev)
    sims = torch.empty(max(capacity, 1), dtype=torch.int32, device=dev) if want_sims else None
    counter = torch.zeros(1, dtype=torch.int64, device=dev)
    check(_lib.load().qrs_pairs_score(ptr(records), records.numel(), ptr(slot_off), ptr(sig), ptr(norm2), sig_id_base,
                                      K, b, ptr(pairs), ptr(sims), capacity, ptr(counter), _stream()),
          "qrs_pairs_score")
    return pairs, sims, counter


def pair_cosine(sig: torch.Tensor, norm2: torch.Tensor, pairs: torch.Tensor) -> torch.Tensor:
    _need_cuda(sig, norm2, pairs)
    sims = torch.empty(pairs.numel(), dtype=torch.int32, device=sig.device)
    check(_lib.load().qrs_pair_cosine(ptr(sig), ptr(norm2), sig.shape[1], ptr(pairs), pairs.numel(), ptr(sims),
                                      _stream()), "qrs_pair_cosine")
    return sims


# ------------------------------------------------------------------------ f-3
def center_rows(rowptr: torch.Tensor, val: torch.Tensor):
    """Mean-centred (truncated, recommender.py:263-267) rating rows: (cval int16[nnz], norm2 int64[rows])."""
    _need_cuda(rowptr, val)
    n_rows = rowptr.numel() - 1
    cval = torch.empty(val.numel(), dtype=torch.int16, device=val.device)
    norm2 = torch.empty(n_rows, dtype=torch.int64, device=val.device)
    check(_lib.load().qrs_center_rows(ptr(rowptr), ptr(val), n_rows, ptr(cval), ptr(norm2), _stream()),
          "qrs_center_rows")
    return cval, norm2


def pair_centered_cosine(rowptr: torch.Tensor, colidx: torch.Tensor, val: torch.Tensor, cval: torch.Tensor,
                         norm2: torch.Tensor, pairs: torch.Tensor, m: int, grouped: bool = False,
                         force_search: bool = False) -> torch.Tensor:
    """int32 max(0, rint(1000 * cos)) of the centred rating rows of each pair (recommender.py:267-269)."""
    _need_cuda(rowptr, colidx, val, cval, norm2, pairs)
    sims = torch.empty(pairs.numel(), dtype=torch.int32, device=pairs.device)
    flags = (_lib.PAIRS_GROUPED if grouped else 0) | (_lib.PAIRS_FORCE_SEARCH if force_search else 0)
    check(_lib.load().qrs_pair_centered_cosine(ptr(rowptr), ptr(colidx), ptr(val), ptr(cval), ptr(norm2), ptr(pairs),
                                               pairs.numel(), m, flags, ptr(sims), _stream()),
          "qrs_pair_centered_cosine")
    return sims


# ------------------------------------------------------------------------ a7/a9
def topk_tables(pairs: torch.Tensor, sims: torch.Tensor, n_pairs: int, row_base: int, n_rows: int, k: int,
                want_milli: bool = False):
    """Neighbour tables (recommender.py:183-210), canonical order: idx int32[n_rows, k] (-1 pad),
    val float64[n_rows, k] (0 pad), cnt int32[n_rows]; with want_milli a 4th tensor int32[n_rows, k]
    holding val * 1000 exactly (the integer form the sparse predictor accumulates)."""
    _need_cuda(pairs, sims)
    lib = _lib.load()
    dev = pairs.device
    adj = torch.empty(max(2 * n_pairs, 1), dtype=torch.int64, device=dev)
    ws = _ws(lib.qrs_topk_workspace_bytes(n_rows), dev)
    idx = torch.empty((n_rows, k), dtype=torch.int32, device=dev)
    val = torch.empty((n_rows, k), dtype=torch.float64, device=dev)
    cnt = torch.empty(n_rows, dtype=torch.int32, device=dev)
    milli = torch.empty((n_rows, k), dtype=torch.int32, device=dev) if want_milli else None
    check(lib.qrs_topk_tables(ptr(pairs), ptr(sims), n_pairs, row_base, n_rows, k, ptr(adj), ptr(ws), ws.numel(),
                              ptr(idx), ptr(val), ptr(milli), ptr(cnt), _stream()), "qrs_topk_tables")
    return (idx, val, cnt, milli) if want_milli else (idx, val, cnt)


# ---------------------------------------------------------------------- a10/a11
Table = Tuple[torch.Tensor, torch.Tensor, torch.Tensor]


def predict_dense(rowptr, colidx, val, Q: int, q_tab: Optional[Table], u_tab: Optional[Table],
                  user_base: int = 0, n_users: Optional[int] = None) -> torch.Tensor:
    """compute_scores' cell loop (recommender.py:311-331) → int32[n_users, Q]."""
    _need_cuda(rowptr, colidx, val)
    U = rowptr.numel() - 1
    n_users = U - user_base if n_users is None else n_users
    out = torch.empty((n_users, Q), dtype=torch.int32, device=rowptr.device)

    def parts(tab):
        if tab is None:
            return None, None, None, 0
        i, v, c = tab
        _need_cuda(i, v, c)
        assert i.dtype == torch.int32 and v.dtype == torch.float64 and c.dtype == torch.int32
        return i, v, c, i.shape[1]

    qi, qv, qc, kq = parts(q_tab)
    ui, uv, uc, ku = parts(u_tab)
    check(_lib.load().qrs_predict_dense(ptr(rowptr), ptr(colidx), ptr(val), U, Q, ptr(qi), ptr(qv), ptr(qc), kq,
                                        ptr(ui), ptr(uv), ptr(uc), ku, user_base, n_users, ptr(out), _stream()),
          "qrs_predict_dense")
    return out


def choose_predict_tile(Q: int, avg_len: float = 0.0, integer_tables: bool = True) -> int:
    """Column window of the sparse predictor.  One WARP accumulates one (user, window) in its slice of shared
    memory, and what the kernel needs most is independent warps: measured at C2 (5 % of 10 000 columns) and C3
    (1 % of 100 000) the time falls with the window width up to the widest window that still leaves the full 20
    warps on an SM (1 024 columns for the integer accumulators) and rises beyond it whatever the row length
    (profiles/r2_predict_variants.md), so that window is taken, rounded so the windows tile Q evenly.  `avg_len`
    is kept for callers that pass it; the choice does not depend on it."""
    lib = _lib.load()
    it = 1 if integer_tables else 0
    top = lib.qrs_predict_warps_per_sm(32, it)
    tile = 32
    while tile + 32 <= 65536 and lib.qrs_predict_warps_per_sm(tile + 32, it) >= top:
        tile += 32
    tile = min(tile, max(32, math.ceil(Q / 32.0) * 32))
    nt = max(1, math.ceil(Q / tile))
    return int(max(32, math.ceil(math.ceil(Q / nt) / 32.0) * 32))


class PeerWorkspace:
    """A device buffer every rank of the process group maps (CUDA IPC): this rank's prepared rating rows live in
    it and the predictors of the other ranks read neighbour rows from it in place over NVLink.  Collective:
    every rank constructs it at the same point.  ptrs[r] = rank r's buffer as mapped here."""

    def __init__(self, nbytes: int, group=None):
        import ctypes
        import torch.distributed as dist
        lib = _lib.load()
        self.nbytes = int(nbytes)
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        p = ctypes.c_void_p()
        check(lib.qrs_peer_alloc(self.nbytes, ctypes.byref(p)), "qrs_peer_alloc")
        self.local = int(p.value)
        handle = (ctypes.c_ubyte * 64)()
        check(lib.qrs_peer_export(self.local, handle), "qrs_peer_export")
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle), group=group)
        self.ptrs = []
        for r, h in enumerate(handles):
            if r == self.rank:
                self.ptrs.append(self.local)
                continue
            q = ctypes.c_void_p()
            buf = (ctypes.c_ubyte * 64).from_buffer_copy(h)
            check(lib.qrs_peer_open(buf, ctypes.byref(q)), "qrs_peer_open")
            self.ptrs.append(int(q.value))
        self.table = (ctypes.c_uint64 * self.world)(*self.ptrs)       # host array handed to qrs_predict_sparse

    def close(self):
        lib = _lib.load()
        for r, q in enumerate(self.ptrs):
            if r != self.rank and q:
                lib.qrs_peer_close(q)
        if self.local:
            lib.qrs_peer_free(self.local)
        self.ptrs, self.local = [], 0


def predict_workspace_bytes(U: int, nnz: int, Q: int, tile: int, capacity: int) -> int:
    return int(_lib.load().qrs_predict_sparse_workspace_bytes(U, nnz, Q, tile, capacity))


class PredictPlan:
    """The row preparation of one ratings CSR for one window width (qrs_predict_prepare): build once, predict any
    number of user ranges from it.  peers: a PeerWorkspace — the preparation is written into this rank's peer
    buffer and predictions may name neighbour rows of any rank by their GLOBAL id (rank * U + row)."""

    def __init__(self, rowptr, colidx, val, Q: int, tile: int, capacity_hint: int = 0, peers: Optional[PeerWorkspace] = None):
        _need_cuda(rowptr, colidx, val)
        if colidx.numel() >= 2 ** 36:
            raise _lib.QrsError("qrs_predict_prepare: nnz < 2^36 per GPU")
        lib = _lib.load()
        self.rowptr, self.colidx, self.val = rowptr, colidx, val
        self.U, self.nnz, self.Q, self.tile = rowptr.numel() - 1, colidx.numel(), int(Q), int(tile)
        self.capacity = int(capacity_hint)
        self.peers = peers
        need = predict_workspace_bytes(self.U, self.nnz, self.Q, self.tile, self.capacity)
        if peers is not None:
            if need > peers.nbytes:
                raise _lib.QrsError("the peer buffer (%d bytes) is too small for this shard's preparation (%d bytes): "
                                    "build the pipeline with a larger peer_margin" % (peers.nbytes, need))
            self.ws, self.ws_ptr, self.ws_bytes = None, peers.local, peers.nbytes
        else:
            self.ws = _ws(need, rowptr.device)
            self.ws_ptr, self.ws_bytes = self.ws.data_ptr(), self.ws.numel()
        check(lib.qrs_predict_prepare(ptr(rowptr), ptr(colidx), ptr(val), self.U, self.nnz, self.Q, self.tile, self.capacity,
                                      self.ws_ptr, self.ws_bytes, _stream()), "qrs_predict_prepare")


class PeerStage:
    """Rows of other ranks that the local users' neighbour table names, copied into a local buffer before the
    predictor runs (qrs_peer_rows_stage).  `cap_steps` 128-byte steps fit; the device word `needed` tells how many
    the table asked for — when it exceeds cap_steps nothing is staged and the predictor reads the rows in place, so
    the caller can size the next step from a readback it does anyway.  Call once every rank has prepared."""

    def __init__(self, plan: "PredictPlan", idx: torch.Tensor, cnt: torch.Tensor, cap_steps: int, keep=None):
        assert plan.peers is not None and idx.dtype == torch.int32 and cnt.dtype == torch.int32
        lib = _lib.load()
        self.n_users, self.ku, self.cap_steps = idx.shape[0], idx.shape[1], int(cap_steps)
        need = int(lib.qrs_peer_rows_stage_bytes(self.n_users, self.ku, plan.Q, plan.tile, self.cap_steps))
        self.buf = keep if keep is not None and keep.numel() >= need else _ws(need, idx.device)
        check(lib.qrs_peer_rows_stage(plan.peers.table, plan.peers.world, plan.peers.rank, plan.U, plan.Q, plan.tile,
                                      ptr(idx), ptr(cnt), self.ku, self.n_users, self.cap_steps,
                                      self.buf.data_ptr(), self.buf.numel(), _stream()), "qrs_peer_rows_stage")
        self.needed = self.buf[:8].view(torch.int64)          # device: steps this table needs


def predict_sparse(rowptr, colidx, val, Q: int, u_tab: Table, user_base: int, n_users: int, tile: int,
                   capacity: int, out_cols: Optional[torch.Tensor] = None, out_preds: Optional[torch.Tensor] = None,
                   u_milli: Optional[torch.Tensor] = None, n_regions: int = 1, cols_mode: int = 0,
                   plan: Optional[PredictPlan] = None, stage: Optional[PeerStage] = None):
    """User-branch predictions in sparse form.  Returns (row_start int64[n_users, nt],
    row_cnt int32[n_users, nt], cols, preds uint8[capacity], cursors int64[n_regions] on device = entries
    requested per region of capacity // n_regions entries, nonzero int64[1] = non-zero predictions).
    u_milli: int32 table with u_val == u_milli / 1000.0 exactly (from topk_tables(want_milli=True)) selects
    the exact-integer accumulation kernel; results are identical.  plan: a PredictPlan of the same CSR and
    window width (prepared rows are reused; its capacity hint must cover `capacity`)."""
    _need_cuda(rowptr, colidx, val, u_milli, *u_tab)
    lib = _lib.load()
    dev = rowptr.device
    U = rowptr.numel() - 1
    nnz = colidx.numel()
    nt = (Q + tile - 1) // tile
    ui, uv, uc = u_tab
    assert u_milli is None or (u_milli.dtype == torch.int32 and u_milli.shape == ui.shape)
    assert ui.dtype == torch.int32 and uv.dtype == torch.float64 and uc.dtype == torch.int32
    row_start = torch.empty((n_users, nt), dtype=torch.int64, device=dev)
    row_cnt = torch.empty((n_users, nt), dtype=torch.int32, device=dev)
    # cols_mode 0: int32 ids; 1: uint16 ids (int16 storage); 2: touched-column bitmap per (user, window), zeroed here
    cols_dtype = torch.int16 if cols_mode == 1 else torch.int32
    if cols_mode == 2:
        cols = torch.zeros(n_users * nt * (tile // 32), dtype=torch.int32, device=dev)
    else:
        cols = out_cols if out_cols is not None else torch.empty(max(capacity, 1), dtype=cols_dtype, device=dev)
    assert cols.dtype == cols_dtype
    preds = out_preds if out_preds is not None else torch.empty(max(capacity, 1), dtype=torch.uint8, device=dev)
    cursor = torch.zeros(n_regions + 1, dtype=torch.int64, device=dev)   # [n_regions] = non-zero count (bitmap mode)
    prepared, peer_tab, n_peers, my_peer = 0, None, 1, 0
    if plan is not None and plan.tile == tile and plan.Q == Q and plan.U == U and plan.nnz == nnz and \
            (plan.capacity >= capacity or plan.peers is not None):
        # the prepared part of the workspace does not move with the capacity
        ws_keep, ws_ptr, ws_bytes, prepared = plan.ws, plan.ws_ptr, plan.ws_bytes, 1
        if plan.peers is not None:
            peer_tab, n_peers, my_peer = plan.peers.table, plan.peers.world, plan.peers.rank
            if predict_workspace_bytes(U, nnz, Q, tile, capacity) > ws_bytes:
                raise _lib.QrsError("the peer buffer is too small for %d predictions per call" % capacity)
    else:
        ws_keep = _ws(lib.qrs_predict_sparse_workspace_bytes(U, nnz, Q, tile, capacity), dev)
        ws_ptr, ws_bytes = ws_keep.data_ptr(), ws_keep.numel()
    check(lib.qrs_predict_sparse(ptr(rowptr), ptr(colidx), ptr(val), U, nnz, Q, ptr(ui), ptr(uv), ptr(u_milli), ptr(uc),
                                 ui.shape[1], user_base, n_users, tile, ws_ptr, ws_bytes, prepared,
                                 peer_tab, n_peers, my_peer,
                                 stage.buf.data_ptr() if stage is not None else None,
                                 stage.n_users if stage is not None else 0,
                                 ptr(row_start), ptr(row_cnt), ptr(cols), cols_mode, ptr(preds), capacity, n_regions,
                                 ptr(cursor), _stream()),
          "qrs_predict_sparse")
    return row_start, row_cnt, cols, preds, cursor[:n_regions], cursor[n_regions:]


# ---------------------------------------------------------------------- f-1: set construction
def answer_sets(table_cm: torch.Tensor, queries: torch.Tensor):
    """compute_shingles' predicate evaluation (recommender.py:76-95) on dictionary codes: table_cm int32[F, drows],
    queries int32[nQ, F] (-1 = no condition) → CSR (rowptr int64[nQ + 1], colidx int32[nnz]) of the answer sets."""
    _need_cuda(table_cm, queries)
    assert table_cm.dtype == torch.int32 and queries.dtype == torch.int32
    F, drows = table_cm.shape
    nq = queries.shape[0]
    assert queries.shape[1] == F
    lib = _lib.load()
    dev = table_cm.device
    rowptr = torch.empty(nq + 1, dtype=torch.int64, device=dev)
    ws = _ws(lib.qrs_answer_sets_workspace_bytes(nq), dev)
    check(lib.qrs_answer_sets_count(ptr(table_cm), drows, F, ptr(queries), nq, ptr(rowptr), ptr(ws), ws.numel(),
                                    _stream()), "qrs_answer_sets_count")
    nnz = int(rowptr[-1].item())
    colidx = torch.empty(max(nnz, 1), dtype=torch.int32, device=dev)
    check(lib.qrs_answer_sets_fill(ptr(table_cm), drows, F, ptr(queries), nq, ptr(rowptr), ptr(colidx), _stream()),
          "qrs_answer_sets_fill")
    return rowptr, colidx[:nnz]


# ---------------------------------------------------------------------- f-2: top-N recommendations
def recommend_topn(row_start, row_cnt, cols, preds, N: int):
    """Per user the N best non-zero predictions (recommender.py:345-374) from predict_sparse's output
    (int32 or uint16 columns): idx int32[n_users, N] (-1 pad), val uint8[n_users, N], cnt int32[n_users]."""
    _need_cuda(row_start, row_cnt, cols, preds)
    n_users, nt = row_start.shape
    mode = 1 if cols.dtype == torch.int16 else 0
    dev = row_start.device
    idx = torch.empty((n_users, N), dtype=torch.int32, device=dev)
    val = torch.empty((n_users, N), dtype=torch.uint8, device=dev)
    cnt = torch.empty(n_users, dtype=torch.int32, device=dev)
    check(_lib.load().qrs_recommend_topn(ptr(row_start), ptr(row_cnt), ptr(cols), mode, ptr(preds), n_users, nt, N,
                                         ptr(idx), ptr(val), ptr(cnt), _stream()), "qrs_recommend_topn")
    return idx, val, cnt


# ---------------------------------------------------------------------- multi-GPU plumbing
def group_pad(sorted_records: torch.Tensor, lg: int, cap: int):
    """Bucket records sorted by their top lg bits (= destination rank) -> (slots int64[2^lg * cap] padded with
    all-ones sentinels, counts int64[2^lg], overflow int32[1]): equal-size slots for an all-to-all without host-side
    split sizes."""
    _need_cuda(sorted_records)
    dev = sorted_records.device
    G = 1 << lg
    out = torch.empty(G * cap, dtype=torch.int64, device=dev)
    counts = torch.zeros(G, dtype=torch.int64, device=dev)
    overflow = torch.zeros(1, dtype=torch.int32, device=dev)
    check(_lib.load().qrs_group_pad(ptr(sorted_records), sorted_records.numel(), lg, cap, ptr(out), ptr(counts),
                                    ptr(overflow), _stream()), "qrs_group_pad")
    return out, counts, overflow


def route_pairs(pairs: torch.Tensor, sims: torch.Tensor, counter: torch.Tensor, rows_per_rank: int, n_ranks: int, cap: int):
    """Scored pairs (the first `counter` entries of pairs / sims) -> (slots int64[n_ranks * cap] of packed words padded
    with sentinels, counts int64[n_ranks], overflow int32[1]); a pair goes to the owner of each of its ends."""
    _need_cuda(pairs, sims, counter)
    dev = pairs.device
    out = torch.empty(n_ranks * cap, dtype=torch.int64, device=dev)
    counts = torch.empty(n_ranks, dtype=torch.int64, device=dev)
    overflow = torch.zeros(1, dtype=torch.int32, device=dev)
    check(_lib.load().qrs_route_pairs(ptr(pairs), ptr(sims), ptr(counter), pairs.numel(), rows_per_rank, n_ranks, cap,
                                      ptr(out), ptr(counts), ptr(overflow), _stream()), "qrs_route_pairs")
    return out, counts, overflow


def unpack_routed(words: torch.Tensor):
    """Received words -> (pairs int64[n], sim_milli int32[n]); sentinel words become all-ones pairs."""
    _need_cuda(words)
    n = words.numel()
    pairs = torch.empty(n, dtype=torch.int64, device=words.device)
    sims = torch.empty(n, dtype=torch.int32, device=words.device)
    check(_lib.load().qrs_unpack_routed(ptr(words), n, ptr(pairs), ptr(sims), _stream()), "qrs_unpack_routed")
    return pairs, sims


# ---------------------------------------------------------------------- f-4: CSV ingest
def csv_line_ends(text: torch.Tensor, max_lines: Optional[int] = None) -> torch.Tensor:
    """Positions of the line ends of a CSV text held on the device (uint8[n]); a last line without '\\n' ends
    at n.  int64[n_lines].  max_lines: an upper bound on the number of lines, if the caller has one (saves the
    sizing pass; a text with more lines falls back to it)."""
    _need_cuda(text)
    assert text.dtype == torch.uint8
    lib = _lib.load()
    n = text.numel()
    dev = text.device
    ws = _ws(lib.qrs_csv_workspace_bytes(n, 0), dev)
    count = torch.zeros(2, dtype=torch.int64, device=dev)
    if n:
        count[1:] = text[-1:]                         # read back together with the line count
    ends, lines = None, -1
    if max_lines is not None and max_lines > 0:
        ends = torch.empty(max_lines + 1, dtype=torch.int64, device=dev)
        check(lib.qrs_csv_line_ends(ptr(text), n, ptr(ends), max_lines, ptr(count), ptr(ws), ws.numel(), _stream()),
              "qrs_csv_line_ends")
        got = count.cpu()
        lines, last = int(got[0]), int(got[1])
        if lines > max_lines:
            ends = None
    if ends is None:
        check(lib.qrs_csv_line_ends(ptr(text), n, None, 0, ptr(count), ptr(ws), ws.numel(), _stream()), "qrs_csv_line_ends")
        got = count.cpu()
        lines, last = int(got[0]), int(got[1])
        ends = torch.empty(lines + 1, dtype=torch.int64, device=dev)
        if lines:
            check(lib.qrs_csv_line_ends(ptr(text), n, ptr(ends), lines, ptr(count), ptr(ws), ws.numel(), _stream()),
                  "qrs_csv_line_ends")
    open_tail = n > 0 and last != 10
    if open_tail:
        ends[lines] = n
    return ends[:lines + (1 if open_tail else 0)]


def csv_ratings(text: torch.Tensor, line_end: torch.Tensor, first_line: int, n_rows: int, Q: int):
    """Rows `first_line ..` of the utility-matrix CSV → (rowptr int64[n_rows + 1], colidx int32[nnz],
    val uint8[nnz], id_end int64[n_rows]); raises QrsError naming the first malformed row."""
    _need_cuda(text, line_end)
    lib = _lib.load()
    dev = text.device
    rowptr = torch.empty(n_rows + 1, dtype=torch.int64, device=dev)
    id_end = torch.empty(max(n_rows, 1), dtype=torch.int64, device=dev)
    err = torch.empty(2, dtype=torch.int32, device=dev)
    ws = _ws(lib.qrs_csv_workspace_bytes(0, n_rows), dev)
    check(lib.qrs_csv_ratings_count(ptr(text), text.numel(), ptr(line_end), first_line, n_rows, Q, ptr(rowptr), ptr(id_end), ptr(err),
                                    ptr(ws), ws.numel(), _stream()), "qrs_csv_ratings_count")
    tail = torch.cat([rowptr[-1:], err.to(torch.int64)]).cpu()
    nnz, n_bad, first_bad = int(tail[0]), int(tail[1]), int(tail[2])
    if n_bad:
        raise _lib.QrsError("utility matrix CSV: %d malformed row(s), first at data row %d (expected %d fields, "
                            "integer ratings 0..255, no quotes)" % (n_bad, first_bad, Q + 1))
    colidx = torch.empty(max(nnz, 1), dtype=torch.int32, device=dev)
    val = torch.empty(max(nnz, 1), dtype=torch.uint8, device=dev)
    check(lib.qrs_csv_ratings_fill(ptr(text), text.numel(), ptr(line_end), first_line, n_rows, Q, ptr(rowptr), ptr(colidx), ptr(val),
                                   _stream()), "qrs_csv_ratings_fill")
    return rowptr, colidx[:nnz], val[:nnz], id_end[:n_rows]


def csv_format(ids: torch.Tensor, ids_off: torch.Tensor, rowptr: torch.Tensor, colidx: torch.Tensor, val: torch.Tensor,
               Q: int, pred=None, zero_as_empty: bool = False) -> torch.Tensor:
    """Rows `user,v,v,...\\r\\n` of the matrix (ratings CSR + optional sparse predictions
    pred = (row_start int64[n, nt], row_cnt int32[n, nt], cols int32|int16, preds uint8)) as uint8 text."""
    _need_cuda(ids, ids_off, rowptr, colidx, val)
    lib = _lib.load()
    dev = rowptr.device
    n_rows = rowptr.numel() - 1
    ps = pc = pcols = pv = None
    nt, mode = 0, 0
    if pred is not None:
        ps, pc, pcols, pv = pred
        _need_cuda(ps, pc, pcols, pv)
        nt = ps.shape[1]
        mode = 1 if pcols.dtype == torch.int16 else 0
    text_off = torch.empty(n_rows + 1, dtype=torch.int64, device=dev)
    ws = _ws(lib.qrs_csv_format_workspace_bytes(n_rows), dev)
    z = 1 if zero_as_empty else 0
    check(lib.qrs_csv_format_count(ptr(ids_off), ptr(rowptr), ptr(val), ptr(ps), ptr(pc), ptr(pv), nt, n_rows, Q, z,
                                   ptr(text_off), ptr(ws), ws.numel(), _stream()), "qrs_csv_format_count")
    total = int(text_off[-1].item())
    text = torch.empty(max(total, 1), dtype=torch.uint8, device=dev)
    check(lib.qrs_csv_format_fill(ptr(ids), ptr(ids_off), ptr(rowptr), ptr(colidx), ptr(val), ptr(ps), ptr(pc), ptr(pcols),
                                  mode, ptr(pv), nt, n_rows, Q, z, ptr(text_off), ptr(text), _stream()),
          "qrs_csv_format_fill")
    return text[:total]
