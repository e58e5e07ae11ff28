"""Multi-GPU invariance check: the sharded pipeline on G ranks must produce the
candidate set, rounded similarities, neighbour tables and predictions of a single-GPU run over the
same matrix (lsh.py:40-54, recommender.py:183-210, 311-331 do not depend on how rows are sharded).

Product against product — no oracle here (the single-GPU path is pinned to the oracle by the GPU
tests and by bench.py's parity check).  Used by `bench.py --gpus N` and `tests/run_multigpu_parity.py`.
"""
from __future__ import annotations

import torch

from . import synth
from .pipeline import LshCfPipeline


def prediction_checksums(out, n_rows: int) -> torch.Tensor:
    """Per-user, order-free checksum of the sparse predictions: sum(col * 131 + pred) and the count,
    packed as checksum * 2^20 + count (int64[n_rows])."""
    dev = out.pred_row_cnt.device
    nt = out.pred_row_start.shape[1]
    cnt = out.pred_row_cnt.reshape(-1).long()
    user_of = torch.repeat_interleave(torch.arange(n_rows * nt, device=dev) // nt, cnt)
    starts = torch.repeat_interleave(out.pred_row_start.reshape(-1), cnt)
    within = torch.arange(user_of.numel(), device=dev) - torch.repeat_interleave(torch.cumsum(cnt, 0) - cnt, cnt)
    pos = starts + within
    cols = out.pred_cols[pos].long()
    if out.pred_cols.dtype == torch.int16:
        cols = cols & 0xFFFF
    cs = torch.zeros(n_rows, dtype=torch.int64, device=dev)
    cs.index_add_(0, user_of, cols * 131 + out.pred_vals[pos].long())
    per_user = torch.zeros(n_rows, dtype=torch.int64, device=dev)
    per_user.index_add_(0, user_of, torch.ones_like(user_of))
    return cs * (1 << 20) + per_user


def _gather_var(t: torch.Tensor, world: int, group=None):
    import torch.distributed as dist
    sizes = [None] * world
    dist.all_gather_object(sizes, int(t.numel()), group=group)
    bufs = [torch.empty(s, dtype=t.dtype, device=t.device) for s in sizes]
    dist.all_gather(bufs, t.contiguous(), group=group)
    return bufs


def multi_rank_invariance(n_per_rank: int, m: int, density: float, K: int, b: int, k: int, perm_tab, dev,
                          seed: int = 77) -> dict:
    """Collective: every rank runs its shard through the distributed pipeline; rank 0 also runs the whole
    matrix on its own GPU and compares.  Returns (on rank 0) {"identical": bool, ...}; {} elsewhere."""
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    rowptr, colidx, val = synth.planted_csr(n_per_rank, m, density, seed=seed, row_base=rank * n_per_rank, device=dev)
    pipe = LshCfPipeline(m=m, K=K, b=b, k=k, perm_tab=perm_tab, device=dev)
    out = pipe.run(rowptr, colidx, val)
    torch.cuda.synchronize()
    pairs_all = _gather_var(out.pairs, world)
    sims_all = _gather_var(out.sims, world)
    idx_all = _gather_var(out.nbr_idx.reshape(-1), world)
    milli_all = _gather_var(torch.round(out.nbr_val.reshape(-1) * 1000).to(torch.int32), world)
    cnt_all = _gather_var(out.nbr_cnt, world)
    cs_all = _gather_var(prediction_checksums(out, n_per_rank), world)
    res = {}
    if rank == 0:
        n = n_per_rank * world
        rp, ci, va = synth.planted_csr(n, m, density, seed=seed, device=dev)
        single = LshCfPipeline(m=m, K=K, b=b, k=k, perm_tab=perm_tab, device=dev, distributed=False).run(rp, ci, va)
        p1, order1 = torch.sort(single.pairs)
        pm, orderm = torch.sort(torch.cat(pairs_all))
        checks = {
            "candidate_pairs": bool(torch.equal(p1, pm)),
            "similarities": bool(p1.numel() == pm.numel() and torch.equal(single.sims[order1], torch.cat(sims_all)[orderm])),
            "neighbour_ids": bool(torch.equal(single.nbr_idx.reshape(-1), torch.cat(idx_all))),
            "neighbour_sims": bool(torch.equal(torch.round(single.nbr_val.reshape(-1) * 1000).to(torch.int32),
                                               torch.cat(milli_all))),
            "neighbour_counts": bool(torch.equal(single.nbr_cnt, torch.cat(cnt_all))),
            "predictions": bool(torch.equal(prediction_checksums(single, n), torch.cat(cs_all))),
        }
        res = {"world": world, "sets": n, "elements": m, "pairs": int(pm.numel()),
               "pairs_scored_per_rank": [int(x.numel()) for x in pairs_all],
               "predictions": int(single.n_predictions), "checks": checks, "identical": all(checks.values())}
    dist.barrier()
    return res
