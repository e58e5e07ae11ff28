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

import os
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
    nbr_milli: torch.Tensor = None         # int32[n_local, k]: 1000 * nbr_val, exact (similarities have 3 decimals)
    nbr_cnt: torch.Tensor = None           # int32[n_local]
    pred_row_start: torch.Tensor = None    # int64[n_local, nt]
    pred_row_cnt: torch.Tensor = None      # int32[n_local, nt]
    pred_cols: torch.Tensor = None         # int32[n_predictions]
    pred_vals: torch.Tensor = None         # uint8[n_predictions]
    pred_chunks: list = None               # per user chunk: {user_lo, n_users, row_start, row_cnt, cols, preds}
    n_predictions: int = 0
    covered_cells: int = 0                 # zero cells of R decided (predicted or proven "missed")
    stage_ms: Dict[str, float] = field(default_factory=dict)
    exchange_bytes: Dict[str, int] = field(default_factory=dict)   # several GPUs: bytes this rank sent per exchange
    peer_rows_staged: bool = False         # several GPUs: remote neighbour rows were copied here before predicting


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
        self._cap = {}                      # learnt buffer capacities: candidate pairs, exchange slots
        self.peer_margin = 1.25             # several GPUs: the peer buffer is sized for this x the first shard's entries
        # QRS_STAGE_PEER_ROWS=1: copy the neighbour rows of other ranks here before predicting instead of reading them
        # in place.  Measured and NOT the default: the copy (SM loads over NVLink: 0.32-0.42 of the link rate) costs
        # more than the predictor saves — C2 shards, N = 2: 6.79 vs 6.55 ms/step, N = 8: 10.14 vs 9.23
        self.stage_peer_rows = os.environ.get("QRS_STAGE_PEER_ROWS", "0")
        self.peer_stage_budget = 8 << 30    # ... unless a step would need more than this many bytes of them
        self._stage_buf = None
        self.predict_chunks = 1             # > 1: predictions are produced (and can be shipped) chunk by chunk
        self.chunk_schedule = "equal"       # "ramp": geometric chunk sizes (see _chunk_bounds)
        self.chunk_growth = 1.5
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
            predict: bool = True, chunk_sink=None, val_ready=None, tables_sink=None) -> PipelineOutput:
        """rowptr/colidx/val: this rank's shard (rows [rank*n_local, (rank+1)*n_local)) of the users×queries
        CSR, on `self.device`.  Every rank must hold the same number of rows.

        Buffer capacities that depend on the data (candidate pairs, exchange slots, prediction slices) are learnt:
        a step runs with the capacities the previous one suggested and reads ONE batch of counters back at its end;
        if any was too small the step is repeated with the corrected value (first call of a new shape only)."""
        for _ in range(5):
            out, ok = self._run_once(rowptr, colidx, val, predict, chunk_sink, val_ready, tables_sink)
            if ok:
                return out
        raise _lib.QrsError("buffer capacities did not settle after 5 attempts")

    def _run_once(self, rowptr, colidx, val, predict, chunk_sink, val_ready, tables_sink=None):
        st, G, rank = self.st, self.world, self.rank
        out = PipelineOutput()
        tm = _Timer(self.profile, self.device.type == "cuda")
        n_local = rowptr.numel() - 1
        row_base = rank * n_local
        n_total = n_local * G
        nnz = colidx.numel()
        dev = rowptr.device
        if G > 1:
            self._check_equal_shards(n_local)
        self._ensure_head(nnz / max(n_local, 1))
        xbytes = {}
        tm.mark("start")

        # a1 + a2/a3: signatures, squared norms and bucket records in one kernel
        sig, norm2, records = st.minhash(rowptr, colidx, self.m, self.perm_tab, self.order_head, b=self.b,
                                         id_base=row_base, want_norm2=True, mode=self.mode)
        out.sig, out.norm2 = sig, norm2
        tm.mark("signatures+banding")

        # global flags / maxima, all-reduced (MAX) once per step: [records overflow, pairs overflow, route overflow,
        # largest record slot, largest route slot]
        flags = torch.zeros(5, dtype=torch.int64, device=dev)
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
            xbytes["signatures"] = (sig.numel() * 4 + norm2.numel() * 8) * (G - 1)
            flat, lg, rec_counts, rec_over, xbytes["records"] = self._exchange_records(records.reshape(-1))
            flags[0] = rec_over[0]
            flags[3] = rec_counts.max()
            tm.mark("exchange:records")
            st.sort_u64(flat, flat.numel(), 1, 32, 64 - lg)
            tm.mark("grouping")
            sig_all, norm2_all = sig_join()
            tm.mark("exchange:signatures")

        # a5 + a8: slots → verified, de-duplicated, scored pairs.  The pair buffer is sized from the previous step
        # (entries beyond the pairs found are sentinels the table builder skips), so no counter is read back here.
        slot_off = st.pair_slots(flat)
        by_ratings = self.similarity == "ratings"
        if by_ratings and G > 1:
            raise _lib.QrsError("similarity='ratings' needs both rows of a pair on one rank: single-rank only")
        if self.similarity not in ("signature", "ratings"):
            raise ValueError("similarity must be 'signature' or 'ratings'")
        cap_pairs = self._cap.get("pairs", 0)
        if by_ratings or cap_pairs == 0:
            cap_pairs = max(int(slot_off[-1].item()), 1)            # exact upper bound (one readback, first step only)
        pairs, sims, counter = st.pairs_score(flat, slot_off, sig_all, norm2_all, self.b, cap_pairs, 0, not by_ratings,
                                              fill_sentinel=True)
        flags[1] = (counter[0] > cap_pairs).to(torch.int64)
        if by_ratings:
            assert val is not None, "similarity='ratings' scores pairs on the ratings themselves"
            n_pairs = int(counter.item())
            pairs = pairs[:n_pairs]
            if val_ready is not None:
                torch.cuda.current_stream().wait_event(val_ready)
            cval, norm2_c = st.center_rows(rowptr, val)
            # equal first ids adjacent: the scoring kernel stages row i once per run
            if n_pairs:
                st.sort_u64(pairs, n_pairs, 1, 32, 32 + max(1, (n_total - 1).bit_length()))
            sims = st.pair_centered_cosine(rowptr, colidx, val, cval, norm2_c, pairs, self.m, grouped=True)
        tm.mark("pairs+cosine")

        if G > 1:
            pairs_t, sims_t, route_counts, route_over, xbytes["pairs"] = self._route_pairs(pairs, sims, counter, n_local)
            flags[2] = route_over[0]
            flags[4] = route_counts.max()
            tm.mark("exchange:pairs")
        else:
            pairs_t, sims_t = pairs, sims

        # a7 + a9
        idx, v, cnt, milli = st.topk_tables(pairs_t, sims_t, pairs_t.numel(), row_base, n_local, self.k, want_milli=True)
        out.nbr_idx, out.nbr_val, out.nbr_cnt, out.nbr_milli = idx, v, cnt, milli
        if tables_sink is not None:
            tables_sink(out)          # the neighbour tables are final here: a caller may ship them while the predictor runs
        tm.mark("topk")

        tail = [slot_off[-1:].to(torch.int64), counter.reshape(1).to(torch.int64)]
        launched = []
        if predict:
            assert val is not None
            if val_ready is not None:          # the ratings were still being uploaded while the LSH stages ran
                torch.cuda.current_stream().wait_event(val_ready)
            tile = self.tile or st.choose_predict_tile(self.m, nnz / max(n_local, 1), self.integer_predict)
            # Users are predicted in `predict_chunks` consecutive chunks so that a caller-supplied sink can
            # ship chunk i to the host while chunk i+1 is computed.  Inside a chunk the output slices are
            # reserved from `regions` cursors; the region capacity is learnt from a first (counting) call and
            # kept for later steps, re-run only if a region would overflow.
            regions = self.pred_regions
            chunks = max(1, min(self.predict_chunks, n_local))
            bounds = self._chunk_bounds(n_local, chunks)
            known = [self._pred_capacity.get((chunks, ci, self.cols_mode), 0) * regions for ci in range(chunks)]
            peers, stage = None, None
            if G > 1:
                # neighbour rows of other ranks are read in place from their prepared buffers: nobody may rewrite
                # its buffer while a peer still predicts from it (barrier 1), nobody may predict before every
                # buffer is written (barrier 2, which also carries the step's overflow flags)
                peers = self._peer_workspace(n_local, nnz, tile)
                self._stream_barrier(dev)
            # the rows are prepared (cut at the window bounds, packed, padded to 128-byte steps) once per step
            plan = st.PredictPlan(rowptr, colidx, val, self.m, tile, capacity_hint=max(known), peers=peers)
            tm.mark("predict:prepare")
            if G > 1:
                self._all_reduce_max(flags)
                tm.mark("exchange:barrier")
                remote = (idx >= 0) & ((idx < row_base) | (idx >= row_base + n_local))
                tail.append(remote.sum().reshape(1).to(torch.int64))
                if self.stage_peer_rows == "1" and hasattr(st, "PeerStage"):
                    # the rows the neighbour tables name on other ranks are copied here in one sweep (NVLink at
                    # bandwidth) instead of being read line by line by the predicting warps (NVLink at latency);
                    # the buffer size is learnt like the other capacities: a step that needs more reads in place
                    stage = st.PeerStage(plan, idx, cnt, self._cap.get("peer_stage", 0), keep=self._stage_buf)
                    self._stage_buf = stage.buf
                    tail.append(stage.needed.reshape(1))
                    tm.mark("exchange:ratings")
            for ci in range(chunks):
                launched.append(self._launch_chunk(st, plan, rowptr, colidx, val, idx, v, cnt, milli, ci, bounds, chunks,
                                                   regions, tile, chunk_sink, stage))
            tail.append((val != 0).sum().reshape(1).to(torch.int64))
        elif G > 1:
            self._all_reduce_max(flags)

        # ---- the step's only readback: every counter at once
        host = torch.cat([flags] + tail + [torch.cat([c["cursors"], c["nz"]]) for c in launched]).cpu().tolist()
        f_rec, f_pairs, f_route, max_rec, max_route = host[:5]
        n_slots, n_found = int(host[5]), int(host[6])
        pos = 5 + len(tail)           # the chunks' counters follow the tail
        ok = True
        if G > 1:
            self._cap["records"] = int(max_rec * 1.1) + 1024
            self._cap["route"] = int(max_route * 1.15) + 4096
            if f_rec or f_route:
                ok = False
        self._cap["pairs"] = int(n_found * 1.15) + 4096
        if f_pairs or n_found > cap_pairs:
            ok = False            # (single rank: the flag is local; several ranks: it was all-reduced)
        if not ok:
            return out, False
        n_pairs = n_found if not by_ratings else pairs.numel()
        out.pairs, out.sims, out.n_slots = pairs[:n_pairs], (sims[:n_pairs] if sims is not None else None), n_slots
        out.exchange_bytes = xbytes
        if predict:
            if G > 1:
                xbytes["ratings(peer reads)"] = int(int(host[7]) * (nnz / max(n_local, 1)) * 4)
                if stage is not None:
                    needed = int(host[8])
                    xbytes["ratings(peer reads)"] = needed * 128
                    out.peer_rows_staged = needed <= stage.cap_steps
                    if needed * 128 <= self.peer_stage_budget:
                        self._cap["peer_stage"] = int(needed * 1.1) + 1024
                    else:                       # (strong-scaled C3: most neighbours are remote — tens of GB)
                        self._cap["peer_stage"] = 0
                        self.stage_peer_rows = "0"
            total = 0
            out.pred_chunks = []
            regions = self.pred_regions
            for c in launched:
                while True:
                    demand = host[pos:pos + regions] if c.get("fresh", True) else None
                    if demand is None:          # the chunk was re-run: read its counters alone
                        again = torch.cat([c["cursors"], c["nz"]]).cpu().tolist()
                        demand, nonzero = again[:regions], int(again[regions])
                    else:
                        nonzero = int(host[pos + regions])
                    need = int(max(demand))
                    if need <= c["rcap"]:
                        break
                    key = (c["chunks"], c["ci"], self.cols_mode)
                    self._pred_capacity[key] = int(need * 1.03) + 256
                    c.update(self._launch_chunk(st, plan, rowptr, colidx, val, idx, v, cnt, milli, c["ci"], c["bounds"],
                                                c["chunks"], regions, c["tile"], chunk_sink, stage))
                    c["fresh"] = False
                pos += regions + 1
                total += nonzero
                out.pred_chunks.append(c["chunk"])
            if len(launched) == 1:
                ch = launched[0]["chunk"]
                out.pred_row_start, out.pred_row_cnt = ch["row_start"], ch["row_cnt"]
                out.pred_cols, out.pred_vals = ch["cols"], ch["preds"]
            out.n_predictions = total
            out.covered_cells = n_local * self.m - int(host[5 + len(tail) - 1])
            tm.mark("predict")
        out.stage_ms = tm.result()
        if self.profile and getattr(self, "_bg_events", None) is not None:
            out.stage_ms["background:signatures"] = self._bg_events[0].elapsed_time(self._bg_events[1])
            self._bg_events = None
        return out, True

    def _chunk_bounds(self, n_local: int, chunks: int):
        """User ranges of the prediction chunks.  "equal": equal parts.  "ramp": each chunk `chunk_growth` times the
        previous one — a caller that ships chunk i while chunk i + 1 is computed (HostRunner) gets its first bytes
        after a small fraction of the prediction time, and since the download is slower than the computation, a
        growth below (download time / prediction time) keeps the copy stream busy from then on."""
        if self.chunk_schedule == "ramp" and chunks > 1:
            w = [self.chunk_growth ** i for i in range(chunks)]
            tot, acc, bounds = sum(w), 0.0, [0]
            for x in w[:-1]:
                acc += x
                bounds.append(int(round(n_local * acc / tot)))
            bounds.append(n_local)
            if all(b2 > b1 for b1, b2 in zip(bounds[:-1], bounds[1:])):
                return bounds
        return [n_local * ci // chunks for ci in range(chunks + 1)]

    def _launch_chunk(self, st, plan, rowptr, colidx, val, idx, v, cnt, milli, ci, bounds, chunks, regions, tile, chunk_sink,
                      stage=None):
        lo, hi = bounds[ci], bounds[ci + 1]
        rcap = self._pred_capacity.get((chunks, ci, self.cols_mode), 0)
        tab = (idx[lo:hi], v[lo:hi], cnt[lo:hi])
        mil = milli[lo:hi] if self.integer_predict else None
        rs, rc, cols, preds, cursors, nzc = st.predict_sparse(rowptr, colidx, val, self.m, tab, lo, hi - lo, tile,
                                                             rcap * regions, u_milli=mil, n_regions=regions,
                                                             cols_mode=self.cols_mode, plan=plan,
                                                             **({"stage": stage} if stage is not None else {}))
        chunk = {"user_lo": lo, "n_users": hi - lo, "row_start": rs, "row_cnt": rc, "cols": cols, "preds": preds,
                 "cols_mode": self.cols_mode, "tile": tile}
        if chunk_sink is not None and rcap > 0:
            chunk_sink(ci, chunk)
        return {"chunk": chunk, "cursors": cursors, "nz": nzc, "rcap": rcap, "ci": ci, "bounds": bounds, "chunks": chunks,
                "tile": tile}

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
            if self.profile:          # the background work's own duration (it overlaps the main stream)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            res = fn()
            if self.profile:
                e1.record()
                self._bg_events = (e0, e1)
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

    def _check_equal_shards(self, n_local: int):
        """row_base = rank * n_local: every rank must hold the same number of rows (checked once per shard size)."""
        import torch.distributed as dist
        if getattr(self, "_shards_ok", None) == n_local:
            return
        sizes = [None] * self.world
        dist.all_gather_object(sizes, int(n_local), group=self.group)
        if any(x != n_local for x in sizes):
            raise _lib.QrsError("every rank must hold the same number of rows, got %s" % sizes)
        self._shards_ok = n_local

    def _stream_barrier(self, dev):
        """Stream-ordered barrier over the ranks: a one-element all-reduce completes on this rank's stream only after
        every rank's stream has reached it (no host synchronisation)."""
        self._all_reduce_max(torch.zeros(1, dtype=torch.int64, device=dev))

    def _all_reduce_max(self, t: torch.Tensor):
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)

    def _peer_workspace(self, n_local: int, nnz: int, tile: int):
        """This rank's peer buffer for the prepared rows (collective, created on first use and kept): sized for
        peer_margin x this shard's entries and for up to min(n_local * m, 64 * nnz) predictions per call."""
        key = (n_local, tile)
        pw = getattr(self, "_peers", None)
        if pw is not None and self._peers_key == key:
            return pw
        if pw is not None:
            pw.close()
        nnz_bound = int(nnz * self.peer_margin) + 4096
        cap_bound = int(min(n_local * self.m, 64 * max(nnz, 1)))
        nbytes = self.st.predict_workspace_bytes(n_local, nnz_bound, self.m, tile, cap_bound)
        self._peers = self.st.PeerWorkspace(nbytes, self.group)
        self._peers_key = key
        return self._peers

    def _exchange_records(self, flat: torch.Tensor):
        """Bucket records go to rank = top log2(G) bits of hash24: one radix pass groups them by destination, the
        groups are padded into G equal slots (sentinel records sort last and pair with nothing) and ONE equal-split
        all-to-all moves them — no split size reaches the host."""
        import torch.distributed as dist
        G = self.world
        lg = G.bit_length() - 1
        if (1 << lg) != G:
            raise _lib.QrsError("the bucket exchange needs a power-of-two world size, got %d" % G)
        n = flat.numel()
        self.st.sort_u64(flat, n, 1, 64 - lg, 64)
        cap = self._cap.get("records") or (int(n / G * 1.5) + 1024)
        slots, counts, over = self.st.group_pad(flat, lg, cap)
        recv = torch.empty_like(slots)
        dist.all_to_all_single(recv, slots, group=self.group)
        return recv, lg, counts, over, slots.numel() * 8 * (G - 1) // G

    def _route_pairs(self, pairs: torch.Tensor, sims: torch.Tensor, counter: torch.Tensor, n_local: int):
        """Each scored pair goes to the owner of i and (if different) the owner of j, as ONE int64
        (i << 37 | j << 11 | sim + 1000; ids < 2^26) appended by a kernel to the destination's fixed-size slot;
        one equal-split all-to-all moves the slots."""
        import torch.distributed as dist
        G = self.world
        if n_local * G >= (1 << 26):
            raise _lib.QrsError("pair routing packs ids into 26 bits; %d sets do not fit" % (n_local * G))
        cap = self._cap.get("route")
        if not cap:
            # first step of a shape: no slot size learnt yet.  A pair sends at most one word to a destination, so the
            # largest pair buffer of any rank bounds every slot; the size must be the same on all ranks (equal splits)
            t = torch.tensor([pairs.numel()], dtype=torch.int64, device=pairs.device)
            self._all_reduce_max(t)
            cap = int(t.item())
        words, counts, over = self.st.route_pairs(pairs, sims, counter, n_local, G, cap)
        recv = torch.empty_like(words)
        dist.all_to_all_single(recv, words, group=self.group)
        p, s = self.st.unpack_routed(recv)
        return p, s, counts, over, words.numel() * 8 * (G - 1) // G


class HostRunner:
    """The call a user of the drop-in makes with HOST data: pinned utility-matrix CSR in; sparse
    predictions + neighbour tables out, in pinned host memory.  Host↔device copies are part of the
    call: the CSR goes up on the compute stream (column ids may be passed as uint16 when the matrix has at most
    65536 queries: half the bytes of the largest input, widened on the device), the neighbour tables come back
    as soon as the top-k stage has written them, predictions come back chunk by chunk on a second stream while
    the next chunk of users is being predicted (touched-column bitmap or uint16 column ids when they fit).
    The similarities travel as the exact integers `nbr_milli` = 1000 * similarity (3 decimals by construction,
    recommender.py:204): `HostRunner.similarities(res)` is the float64 table.
    Buffers are allocated once and reused across calls."""

    def __init__(self, pipe: LshCfPipeline, chunks: int = 8, cols_format: str = "auto", schedule: str = "ramp"):
        """cols_format: "i32" | "u16" | "bitmap" | "auto" (u16 when the column ids fit, switching to the
        touched-column bitmap once a call has shown that more than 1/16 of the cells get a prediction —
        the point where m/8 bitmap bytes per user beat 2 bytes per prediction)."""
        self.pipe = pipe
        pipe.predict_chunks = chunks
        pipe.chunk_schedule = schedule
        self.cols_format = cols_format
        fits16 = pipe.m <= 65536
        pipe.cols_mode = {"i32": 0, "u16": 1, "bitmap": 2, "auto": 1 if fits16 else 0}[cols_format]
        self._dev = {}
        self._host = {}
        self.copy_stream = torch.cuda.Stream(device=pipe.device)
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    @staticmethod
    def similarities(res) -> torch.Tensor:
        """float64 similarity table of a result (the same doubles the device holds: milli / 1000.0)."""
        return res["nbr_milli"].to(torch.float64) / 1000.0

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
            if name == "colidx" and t.dtype in (torch.uint16, torch.int16):      # compact wire format
                if self.pipe.m > 65536:
                    raise _lib.QrsError("16-bit column ids need a matrix of at most 65536 queries")
                wide = self._dev.get("colidx32")
                if wide is None or wide.numel() < t.numel():
                    wide = self._dev["colidx32"] = torch.empty(t.numel(), dtype=torch.int32, device=self.pipe.device)
                d = self.pipe.st.widen_cols_u16(d, wide)
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
        shipped = {}

        def sink(ci, chunk):
            ready = torch.cuda.Event()
            ready.record(compute)
            host_chunk = {"user_lo": chunk["user_lo"], "n_users": chunk["n_users"], "cols_mode": chunk["cols_mode"],
                          "tile": chunk["tile"]}
            with torch.cuda.stream(self.copy_stream):
                self.copy_stream.wait_event(ready)
                for key in ("row_start", "row_cnt", "cols", "preds"):
                    self._ship("c%d_%s" % (ci, key), chunk[key], host_chunk, key)
            shipped[ci] = host_chunk          # a chunk that had to be re-run replaces its first shipment

        def tables(out):
            ready = torch.cuda.Event()
            ready.record(compute)
            with torch.cuda.stream(self.copy_stream):
                self.copy_stream.wait_event(ready)
                for name in ("nbr_idx", "nbr_milli", "nbr_cnt"):
                    self._ship(name, getattr(out, name), res)

        out = self.pipe.run(*dev_in, predict=True, chunk_sink=sink, val_ready=val_ready, tables_sink=tables)
        self.copy_stream.synchronize()
        compute.synchronize()
        res["chunks"] = [shipped[ci] for ci in sorted(shipped)]
        res["n_predictions"] = out.n_predictions
        res["covered_cells"] = out.covered_cells
        if self.cols_format == "auto" and self.pipe.cols_mode != 2:
            n_local = rowptr_h.numel() - 1
            if out.n_predictions * 16 > n_local * self.pipe.m:
                self.pipe.cols_mode = 2          # later calls ship bitmap + one byte per touched column
        return res
