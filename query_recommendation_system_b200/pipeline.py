"""Device-resident LSH → top-k CF pipeline for Binding B (set = user row of the utility matrix,
element = rated query): signatures → bucket records → grouping → candidate pairs +
rounded cosine → neighbour tables → similarity-weighted predictions, i.e. the chain
recommender.py:149 → 171-176 → 183-210 → 311-331 without ever leaving the GPU.

Multi-GPU (one process per GPU, torch.distributed): rows are sharded contiguously; the
signature matrix is all-gathered, bucket records are exchanged with an all-to-all keyed by the
top bits of the bucket hash, pairs are scored where their bucket lands, scored pairs are sent
to the owners of their two ends, and only the neighbour rows that live on other ranks are fetched
(request/response all-to-all) for the prediction gathers.
All arithmetic is in `stages` (CUDA kernels behind the C ABI); torch is plumbing.
"""
from __future__ import annotations

import time
from dataclasses import dataclass, field
from typing import Dict, Optional

import numpy as np
import torch

from . import _lib
from . import stages as cuda_stages


@dataclass
class PipelineOutput:
    sig: torch.Tensor = None               # int32[n_local, K]
    norm2: torch.Tensor = None             # int64[n_local]
    pairs: torch.Tensor = None             # int64[P] packed (i << 32 | j), pairs scored on this rank
    sims: torch.Tensor = None              # int32[P] rint(1000 cos)
    n_slots: int = 0                       # pre-dedup (pair, band) coincidences examined on this rank
    nbr_idx: torch.Tensor = None           # int32[n_local, k]
    nbr_val: torch.Tensor = None           # float64[n_local, k]
    nbr_cnt: torch.Tensor = None           # int32[n_local]
    pred_row_start: torch.Tensor = None    # int64[n_local, nt]
    pred_row_cnt: torch.Tensor = None      # int32[n_local, nt]
    pred_cols: torch.Tensor = None         # int32[n_predictions]
    pred_vals: torch.Tensor = None         # uint8[n_predictions]
    pred_chunks: list = None               # per user chunk: {user_lo, n_users, row_start, row_cnt, cols, preds}
    n_predictions: int = 0
    covered_cells: int = 0                 # zero cells of R decided (predicted or proven "missed")
    stage_ms: Dict[str, float] = field(default_factory=dict)


class _Timer:
    """CUDA-event stage timer on the current stream (wall clock when the tensors are on the CPU,
    which only happens under the test double)."""

    def __init__(self, enabled: bool, cuda: bool):
        self.enabled, self.cuda = enabled, cuda
        self.marks = []

    def mark(self, name: str):
        if not self.enabled:
            return
        if self.cuda:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self.marks.append((name, ev))
        else:
            self.marks.append((name, time.perf_counter()))

    def result(self) -> Dict[str, float]:
        out = {}
        if not self.enabled or len(self.marks) < 2:
            return out
        if self.cuda:
            torch.cuda.synchronize()
        for (_, a), (name, b) in zip(self.marks[:-1], self.marks[1:]):
            out[name] = out.get(name, 0.0) + (a.elapsed_time(b) if self.cuda else (b - a) * 1e3)
        return out


class LshCfPipeline:
    def __init__(self, m: int, K: int, b: int, k: int, perm_tab, device="cuda", tile: Optional[int] = None,
                 minhash_mode: int = _lib.MINHASH_AUTO, head_ranks: Optional[int] = None,
                 group=None, stages=None, profile: bool = False, distributed: bool = True):
        """perm_tab: element-major permutation table int32[m, K] (numpy or tensor) — the K successive
        np.random.permutation(m) draws of recommender.py:120, `oracle`-independent: the caller makes it."""
        if K % b != 0:
            raise AssertionError("K=%d is not a multiple of b=%d" % (K, b))      # lsh.py:20
        self.m, self.K, self.b, self.k = int(m), int(K), int(b), int(k)
        self.st = stages if stages is not None else cuda_stages
        self.device = torch.device(device)
        self.group = group
        self.distributed = distributed      # False: ignore an initialised process group (single-rank run)
        self.profile = profile
        self.integer_predict = True         # exact-integer accumulation in the predictor (identical results)
        # output-slice cursors: one per SM, so that every region serves the same number of CTAs of the
        # prediction kernel (its grid is a multiple of the SM count) and the regions fill evenly
        self.pred_regions = 1
        if self.device.type == "cuda":
            self.pred_regions = max(1, _lib.load().qrs_device_sm_count(self.device.index or 0))
        self.tile = tile
        self.mode = minhash_mode
        if isinstance(perm_tab, np.ndarray):
            perm_tab = torch.from_numpy(np.ascontiguousarray(perm_tab.astype(np.int32, copy=False)))
        self.perm_tab = perm_tab.to(self.device, torch.int32).contiguous()
        assert tuple(self.perm_tab.shape) == (self.m, self.K)
        self._head_ranks = head_ranks
        self.order_head = None
        self._head_for = None
        self._pred_capacity = {}
        self.predict_chunks = 1             # > 1: predictions are produced (and can be shipped) chunk by chunk
        self.cols_mode = 0                  # prediction columns as int32 ids (0), uint16 ids (1) or a bitmap (2)
        # how a candidate pair is scored: "signature" = rounded cosine of the MinHash signatures
        # (recommender.py:203-204, the query branch); "ratings" = rounded, clamped cosine of the mean-centred
        # rating rows (recommender.py:263-269, the user branch: LSH candidates stand in for PCA + BIRCH clusters)
        self.similarity = "signature"

    # ------------------------------------------------------------------ distributed helpers
    @property
    def world(self) -> int:
        import torch.distributed as dist
        if not self.distributed:
            return 1
        return dist.get_world_size(self.group) if (dist.is_available() and dist.is_initialized()) else 1

    @property
    def rank(self) -> int:
        import torch.distributed as dist
        if not self.distributed:
            return 0
        return dist.get_rank(self.group) if (dist.is_available() and dist.is_initialized()) else 0

    def _ensure_head(self, avg_len: float):
        R = self._head_ranks if self._head_ranks is not None else self.st.choose_head_ranks(self.m, avg_len)
        if self._head_for != R:
            self.order_head = self.st.perm_head(self.perm_tab, R) if R > 0 else None
            self._head_for = R

    # ------------------------------------------------------------------ the path
    def run(self, rowptr: torch.Tensor, colidx: torch.Tensor, val: Optional[torch.Tensor],
            predict: bool = True, chunk_sink=None, val_ready=None) -> PipelineOutput:
        """rowptr/colidx/val: this rank's shard (rows [rank*n_local, (rank+1)*n_local)) of the users×queries
        CSR, on `self.device`.  Every rank must hold the same number of rows."""
        st, G, rank = self.st, self.world, self.rank
        out = PipelineOutput()
        tm = _Timer(self.profile, self.device.type == "cuda")
        n_local = rowptr.numel() - 1
        row_base = rank * n_local
        n_total = n_local * G
        nnz = colidx.numel()
        self._ensure_head(nnz / max(n_local, 1))
        tm.mark("start")

        lens_join = None
        if G > 1 and predict:
            # row lengths of every shard (4 bytes per row): lets the remote-row exchange size its buffers locally
            lens_join = self._background(
                lambda: (self._all_gather_rows((rowptr[1:] - rowptr[:-1]).contiguous(), self._bg_group()),), (rowptr,))

        # a1 + a2/a3: signatures, squared norms and bucket records in one kernel
        sig, norm2, records = st.minhash(rowptr, colidx, self.m, self.perm_tab, self.order_head, b=self.b,
                                         id_base=row_base, want_norm2=True, mode=self.mode)
        out.sig, out.norm2 = sig, norm2
        tm.mark("signatures+banding")

        if G == 1:
            flat = records.reshape(-1)
            st.sort_u64(flat, n_local, self.b, 40, 64)
            sig_all, norm2_all = sig, norm2
            tm.mark("grouping")
        else:
            # the gathered signatures are first needed by the pair scoring: gather them on a second stream and
            # communicator while the bucket records are exchanged and sorted
            sig_join = self._background(lambda: (self._all_gather_rows(sig, self._bg_group()),
                                                 self._all_gather_rows(norm2, self._bg_group())), (sig, norm2))
            flat, lg = self._exchange_records(records.reshape(-1))
            tm.mark("exchange:records")
            st.sort_u64(flat, flat.numel(), 1, 32, 64 - lg)
            tm.mark("grouping")
            sig_all, norm2_all = sig_join()
            tm.mark("exchange:signatures")

        # a5 + a8: slots → verified, de-duplicated, scored pairs
        slot_off = st.pair_slots(flat)
        n_slots = int(slot_off[-1].item())
        by_ratings = self.similarity == "ratings"
        if by_ratings and G > 1:
            raise _lib.QrsError("similarity='ratings' needs both rows of a pair on one rank: single-rank only")
        if self.similarity not in ("signature", "ratings"):
            raise ValueError("similarity must be 'signature' or 'ratings'")
        pairs, sims, counter = st.pairs_score(flat, slot_off, sig_all, norm2_all, self.b, n_slots, 0, not by_ratings)
        n_pairs = int(counter.item())
        pairs = pairs[:n_pairs]
        if by_ratings:
            assert val is not None, "similarity='ratings' scores pairs on the ratings themselves"
            if val_ready is not None:
                torch.cuda.current_stream().wait_event(val_ready)
            cval, norm2_c = st.center_rows(rowptr, val)
            # equal first ids adjacent: the scoring kernel stages row i once per run
            if n_pairs:
                st.sort_u64(pairs, n_pairs, 1, 32, 32 + max(1, (n_total - 1).bit_length()))
            sims = st.pair_centered_cosine(rowptr, colidx, val, cval, norm2_c, pairs, self.m, grouped=True)
        else:
            sims = sims[:n_pairs]
        out.pairs, out.sims, out.n_slots = pairs, sims, n_slots
        tm.mark("pairs+cosine")

        if G > 1:
            pairs, sims = self._route_pairs(pairs, sims, n_local)
            n_pairs = pairs.numel()
            tm.mark("exchange:pairs")

        # a7 + a9
        idx, v, cnt, milli = st.topk_tables(pairs, sims, n_pairs, row_base, n_local, self.k, want_milli=True)
        out.nbr_idx, out.nbr_val, out.nbr_cnt = idx, v, cnt
        tm.mark("topk")

        if predict:
            assert val is not None
            if val_ready is not None:          # the ratings were still being uploaded while the LSH stages ran
                torch.cuda.current_stream().wait_event(val_ready)
            pred_base = row_base
            if G > 1:
                # only the neighbour rows that live on other ranks are fetched (deduplicated), appended to the local
                # CSR, and the neighbour ids are re-pointed at the extended row numbering
                rowptr_all, colidx_all, val_all, idx = self._fetch_remote_rows(rowptr, colidx, val, idx, cnt,
                                                                               n_local, row_base, lens_join()[0])
                pred_base = 0
                tm.mark("exchange:ratings")
            else:
                rowptr_all, colidx_all, val_all = rowptr, colidx, val
            tile = self.tile or st.choose_predict_tile(self.m, nnz / max(n_local, 1), self.integer_predict)
            # Users are predicted in `predict_chunks` consecutive chunks so that a caller-supplied sink can
            # ship chunk i to the host while chunk i+1 is computed.  Inside a chunk the output slices are
            # reserved from `regions` cursors; the region capacity is learnt from a first (counting) call and
            # kept for later steps, re-run only if a region would overflow.
            regions = self.pred_regions
            chunks = max(1, min(self.predict_chunks, n_local))
            bounds = [n_local * ci // chunks for ci in range(chunks + 1)]
            total = 0
            out.pred_chunks = []
            # the rows are prepared (cut at the window bounds, packed, padded to 128-byte steps) once per step
            known = [self._pred_capacity.get((chunks, ci, self.cols_mode), 0) * regions for ci in range(chunks)]
            plan = st.PredictPlan(rowptr_all, colidx_all, val_all, self.m, tile, capacity_hint=max(known))
            tm.mark("predict:prepare")
            for ci in range(chunks):
                lo, hi = bounds[ci], bounds[ci + 1]
                rcap = self._pred_capacity.get((chunks, ci, self.cols_mode), 0)
                tab = (idx[lo:hi], v[lo:hi], cnt[lo:hi])
                mil = milli[lo:hi] if self.integer_predict else None
                while True:
                    rs, rc, cols, preds, cursors, nzc = st.predict_sparse(rowptr_all, colidx_all, val_all, self.m, tab,
                                                                     pred_base + lo, hi - lo, tile, rcap * regions,
                                                                     u_milli=mil, n_regions=regions,
                                                                     cols_mode=self.cols_mode, plan=plan)
                    demand = torch.cat([cursors, nzc]).cpu()
                    nonzero = int(demand[-1].item())
                    demand = demand[:regions]
                    need = int(demand.max().item())
                    if need <= rcap:
                        break
                    rcap = self._pred_capacity[(chunks, ci, self.cols_mode)] = int(need * 1.03) + 256
                total += nonzero
                chunk = {"user_lo": lo, "n_users": hi - lo, "row_start": rs, "row_cnt": rc, "cols": cols, "preds": preds,
                         "cols_mode": self.cols_mode, "tile": tile}
                out.pred_chunks.append(chunk)
                if chunk_sink is not None:
                    chunk_sink(ci, chunk)
            if chunks == 1:
                out.pred_row_start, out.pred_row_cnt = rs, rc
                out.pred_cols, out.pred_vals = cols, preds
            out.n_predictions = total
            rated = int((val != 0).sum().item()) if val.numel() else 0
            out.covered_cells = n_local * self.m - rated
            tm.mark("predict")
        out.stage_ms = tm.result()
        return out

    def recommend_top_n(self, out: PipelineOutput, N: int = 10):
        """What Recommender.top_k_queries serves (recommender.py:345-374), for every user of this rank at once:
        the N best non-zero predictions among the cells just predicted, without a dense users x queries matrix.
        Returns (idx int32[n_local, N] (-1 pad), val uint8[n_local, N], cnt int32[n_local])."""
        if out.pred_row_start is None:
            raise _lib.QrsError("recommend_top_n needs the single-chunk device output of run() (int32/uint16 columns)")
        return self.st.recommend_topn(out.pred_row_start, out.pred_row_cnt, out.pred_cols, out.pred_vals, N)

    # ------------------------------------------------------------------ exchanges (plumbing)
    def _bg_group(self):
        """A second communicator for background collectives (collectives of one communicator run in order)."""
        import torch.distributed as dist
        if getattr(self, "_bg", None) is None:
            self._bg = dist.new_group(ranks=list(range(self.world))) if self.group is None else self.group
        return self._bg

    def _background(self, fn, inputs):
        """Run `fn` on a side CUDA stream (inline on the CPU test double); returns a getter that joins it."""
        if self.device.type != "cuda":
            res = fn()
            return lambda: res
        if getattr(self, "_side", None) is None:
            self._side = torch.cuda.Stream(device=self.device)
        main = torch.cuda.current_stream()
        self._side.wait_stream(main)
        with torch.cuda.stream(self._side):
            res = fn()
        for t in inputs:
            t.record_stream(self._side)

        def join():
            torch.cuda.current_stream().wait_stream(self._side)
            for t in res:
                t.record_stream(torch.cuda.current_stream())
            return res
        return join

    def _all_gather_rows(self, t: torch.Tensor, group=None) -> torch.Tensor:
        import torch.distributed as dist
        G = self.world
        outt = torch.empty((t.shape[0] * G,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(outt, t.contiguous(), group=group if group is not None else self.group)
        return outt

    def _all_to_all_var(self, send: torch.Tensor, send_counts, return_counts: bool = False, recv_counts=None):
        """send is grouped by destination rank; send_counts: G host ints.  The receive split is exchanged first
        unless the caller already knows it (recv_counts)."""
        import torch.distributed as dist
        send_l = [int(x) for x in (send_counts.tolist() if hasattr(send_counts, "tolist") else send_counts)]
        if recv_counts is None:
            sc = torch.tensor(send_l, dtype=torch.int64, device=send.device)
            rcv = torch.empty_like(sc)
            dist.all_to_all_single(rcv, sc, group=self.group)
            rcv_l = rcv.cpu().tolist()
        else:
            rcv_l = [int(x) for x in recv_counts]
        recv = torch.empty((int(sum(rcv_l)),) + tuple(send.shape[1:]), dtype=send.dtype, device=send.device)
        dist.all_to_all_single(recv, send, output_split_sizes=rcv_l, input_split_sizes=send_l, group=self.group)
        return (recv, rcv_l) if return_counts else recv

    def _exchange_records(self, flat: torch.Tensor):
        """Bucket records go to rank = top log2(G) bits of hash24: one radix pass groups them by
        destination, the split points come from a searchsorted on those bits."""
        G = self.world
        lg = G.bit_length() - 1
        if (1 << lg) != G:
            raise _lib.QrsError("the bucket exchange needs a power-of-two world size, got %d" % G)
        self.st.sort_u64(flat, flat.numel(), 1, 64 - lg, 64)
        dest = (flat >> (64 - lg)) & (G - 1)
        bounds = torch.searchsorted(dest, torch.arange(G + 1, device=flat.device, dtype=dest.dtype))
        counts = (bounds[1:] - bounds[:-1]).cpu()
        return self._all_to_all_var(flat, counts), lg

    def _route_pairs(self, pairs: torch.Tensor, sims: torch.Tensor, n_local: int):
        """Each scored pair goes to the owner of i and (if different) the owner of j.  Pair and similarity
        travel as ONE int64 (i << 37 | j << 11 | sim + 1000; ids < 2^26) so a single all-to-all moves both."""
        G = self.world
        if n_local * G >= (1 << 26):
            raise _lib.QrsError("pair routing packs ids into 26 bits; %d sets do not fit" % (n_local * G))
        i, j = pairs >> 32, pairs & 0xFFFFFFFF
        oi = torch.div(i, n_local, rounding_mode="floor")
        oj = torch.div(j, n_local, rounding_mode="floor")
        rec = (i << 37) | (j << 11) | (sims.to(torch.int64) + 1000)
        cross = oj != oi
        rec = torch.cat([rec, rec[cross]])
        dest, order = torch.sort(torch.cat([oi, oj[cross]]), stable=True)
        counts = torch.bincount(dest, minlength=G).cpu()
        got = self._all_to_all_var(rec[order], counts)
        return ((got >> 37) << 32) | ((got >> 11) & 0x3FFFFFF), ((got & 0x7FF) - 1000).to(torch.int32)

    def _fetch_remote_rows(self, rowptr, colidx, val, idx, cnt, n_local: int, row_base: int, lens_all=None):
        """Request/response exchange of the ratings rows this rank's neighbour tables point at on other ranks.
        Returns the extended CSR (local rows first, fetched rows after, in ascending global id) and the
        neighbour table re-indexed into it.  lens_all: row lengths of ALL ranks' rows (gathered in the
        background at the start of the step), so the requester sizes its receive buffers without a round trip."""
        import torch.distributed as dist
        G = self.world
        dev = idx.device
        k = idx.shape[1]
        live = torch.arange(k, device=dev)[None, :] < cnt[:, None]
        ids = idx.to(torch.int64)
        is_remote = live & ((ids < row_base) | (ids >= row_base + n_local))
        need = torch.unique(ids[is_remote])                                # sorted ⇒ grouped by owner
        owner = torch.div(need, n_local, rounding_mode="floor")
        # one all-gather of the G x G request-count matrix tells every rank who asks whom for how many rows
        mine = torch.bincount(owner, minlength=G)
        counts = torch.empty(G * G, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(counts, mine, group=self.group)
        counts = counts.view(G, G).cpu()
        if int(counts.sum().item()) == 0:                                  # every neighbour of every rank is local
            new_idx = torch.where(live, ids - row_base, torch.full_like(ids, -1)).to(torch.int32)
            return rowptr, colidx, val, new_idx.contiguous()
        need_counts = counts[self.rank].tolist()                           # rows I ask of each owner
        req_counts = counts[:, self.rank].tolist()                         # rows each rank asks of me
        req = self._all_to_all_var(need, need_counts, recv_counts=req_counts)
        # serve the requests: the rows packed by a gather kernel; both sides know every row length
        r = req - row_base
        lens = rowptr[r + 1] - rowptr[r]
        csum = torch.cat([torch.zeros(1, dtype=torch.int64, device=dev), torch.cumsum(lens, 0)])
        lens_got = lens_all[need] if lens_all is not None else \
            self._all_to_all_var(lens.contiguous(), req_counts, recv_counts=need_counts)
        gsum = torch.cat([torch.zeros(1, dtype=torch.int64, device=dev), torch.cumsum(lens_got, 0)])
        rends = torch.cumsum(torch.tensor(req_counts, dtype=torch.int64, device=dev), 0)
        nends = torch.cumsum(torch.tensor(need_counts, dtype=torch.int64, device=dev), 0)
        splits = torch.cat([csum[rends], gsum[nends]]).cpu()               # ONE readback for both split tables
        ent_send = torch.diff(splits[:G], prepend=torch.zeros(1, dtype=torch.int64)).tolist()
        ent_recv = torch.diff(splits[G:], prepend=torch.zeros(1, dtype=torch.int64)).tolist()
        tot = int(sum(ent_send))
        col_send, val_send = self.st.gather_rows(rowptr, colidx, val, r.contiguous(), csum, tot)
        if self.m <= 65536:      # column ids fit 16 bits: ship them as such (3 instead of 5 bytes per entry)
            got = self._all_to_all_var(col_send.to(torch.int16).contiguous().view(torch.uint8),
                                       [2 * x for x in ent_send], recv_counts=[2 * x for x in ent_recv])
            col_got = got.view(torch.int16).to(torch.int32) & 0xFFFF
        else:
            col_got = self._all_to_all_var(col_send.contiguous(), ent_send, recv_counts=ent_recv)
        val_got = self._all_to_all_var(val_send.contiguous(), ent_send, recv_counts=ent_recv)
        # extended CSR and re-indexed table
        rp = torch.cat([rowptr, rowptr[-1] + torch.cumsum(lens_got, 0)])
        col_all = torch.cat([colidx, col_got])
        val_all = torch.cat([val, val_got])
        local_ix = ids - row_base
        remote_ix = n_local + torch.searchsorted(need, ids.clamp(min=0))
        new_idx = torch.where(is_remote, remote_ix, local_ix)
        new_idx = torch.where(live, new_idx, torch.full_like(new_idx, -1)).to(torch.int32)
        return rp, col_all, val_all, new_idx.contiguous()


class HostRunner:
    """The call a user of the drop-in makes with HOST data: pinned utility-matrix CSR in; sparse
    predictions + neighbour tables out, in pinned host memory.  Host↔device copies are part of the
    call: the CSR goes up on the compute stream, predictions come back chunk by chunk on a second
    stream while the next chunk of users is being predicted (column ids as uint16 when they fit).
    Buffers are allocated once and reused across calls."""

    def __init__(self, pipe: LshCfPipeline, chunks: int = 4, cols_format: str = "auto"):
        """cols_format: "i32" | "u16" | "bitmap" | "auto" (u16 when the column ids fit, switching to the
        touched-column bitmap once a call has shown that more than 1/16 of the cells get a prediction —
        the point where m/8 bitmap bytes per user beat 2 bytes per prediction)."""
        self.pipe = pipe
        pipe.predict_chunks = chunks
        self.cols_format = cols_format
        fits16 = pipe.m <= 65536
        pipe.cols_mode = {"i32": 0, "u16": 1, "bitmap": 2, "auto": 1 if fits16 else 0}[cols_format]
        self._dev = {}
        self._host = {}
        self.copy_stream = torch.cuda.Stream(device=pipe.device)
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def _dev_like(self, name, t):
        buf = self._dev.get(name)
        if buf is None or buf.numel() < t.numel() or buf.dtype != t.dtype:
            buf = self._dev[name] = torch.empty(t.numel(), dtype=t.dtype, device=self.pipe.device)
        return buf[:t.numel()]

    def _host_like(self, name, t):
        buf = self._host.get(name)
        if buf is None or buf.numel() < t.numel() or buf.dtype != t.dtype:
            buf = self._host[name] = torch.empty(int(t.numel() * 1.05) + 16, dtype=t.dtype, pin_memory=True)
        return buf[:t.numel()]

    def _ship(self, name, t, res, key=None):
        """D2H of one device tensor into its pinned twin on the copy stream (after the producing kernels)."""
        t = t.reshape(-1)
        h = self._host_like(name, t)
        h.copy_(t, non_blocking=True)
        t.record_stream(self.copy_stream)
        self.d2h_bytes += t.numel() * t.element_size()
        res[key or name] = h

    def __call__(self, rowptr_h: torch.Tensor, colidx_h: torch.Tensor, val_h: torch.Tensor):
        self.h2d_bytes = self.d2h_bytes = 0
        compute = torch.cuda.current_stream()
        dev_in = []
        for name, t in (("rowptr", rowptr_h), ("colidx", colidx_h)):
            d = self._dev_like(name, t)
            d.copy_(t, non_blocking=True)
            self.h2d_bytes += t.numel() * t.element_size()
            dev_in.append(d)
        # the rating values are needed only by the prediction stage: upload them on the copy stream while the
        # signature / LSH stages run
        d_val = self._dev_like("val", val_h)
        self.copy_stream.wait_stream(compute)          # the buffer may still be read by the previous call
        with torch.cuda.stream(self.copy_stream):
            d_val.copy_(val_h, non_blocking=True)
            val_ready = torch.cuda.Event()
            val_ready.record(self.copy_stream)
        self.h2d_bytes += val_h.numel() * val_h.element_size()
        dev_in.append(d_val)
        res = {"chunks": []}

        def sink(ci, chunk):
            ready = torch.cuda.Event()
            ready.record(compute)
            host_chunk = {"user_lo": chunk["user_lo"], "n_users": chunk["n_users"], "cols_mode": chunk["cols_mode"],
                          "tile": chunk["tile"]}
            with torch.cuda.stream(self.copy_stream):
                self.copy_stream.wait_event(ready)
                for key in ("row_start", "row_cnt", "cols", "preds"):
                    self._ship("c%d_%s" % (ci, key), chunk[key], host_chunk, key)
            res["chunks"].append(host_chunk)

        out = self.pipe.run(*dev_in, predict=True, chunk_sink=sink, val_ready=val_ready)
        ready = torch.cuda.Event()
        ready.record(compute)
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(ready)
            for name in ("nbr_idx", "nbr_val", "nbr_cnt"):
                self._ship(name, getattr(out, name), res)
        self.copy_stream.synchronize()
        compute.synchronize()
        res["n_predictions"] = out.n_predictions
        res["covered_cells"] = out.covered_cells
        if self.cols_format == "auto" and self.pipe.cols_mode != 2:
            n_local = rowptr_h.numel() - 1
            if out.n_predictions * 16 > n_local * self.pipe.m:
                self.pipe.cols_mode = 2          # later calls ship bitmap + one byte per touched column
        return res
