#!/usr/bin/env python
"""Multi-GPU invariance check on real GPUs: run under torchrun with N ranks,

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29517 tests/run_multigpu_parity.py [--n-per-rank 20000]

Every rank runs the sharded pipeline (NCCL exchanges); rank 0 additionally runs the same
matrix on one GPU and checks that the union of scored pairs, the neighbour tables and the
predictions of every shard are identical.  Prints one JSON line."""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-per-rank", type=int, default=20000)
    ap.add_argument("--m", type=int, default=4000)
    args = ap.parse_args()
    from query_recommendation_system_b200 import synth
    from query_recommendation_system_b200.pipeline import LshCfPipeline
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n, m, K, b, k = args.n_per_rank, args.m, 128, 32, 24
    np.random.seed(7)
    tab = np.stack([np.random.permutation(m) for _ in range(K)], axis=1).astype(np.int32)
    rowptr, colidx, val = synth.planted_csr(n, m, 0.05, seed=77, row_base=rank * n, device=dev)
    pipe = LshCfPipeline(m=m, K=K, b=b, k=k, perm_tab=tab, device=dev)
    out = pipe.run(rowptr, colidx, val)
    torch.cuda.synchronize()

    def gather_var(t):
        sizes = [None] * world
        dist.all_gather_object(sizes, int(t.numel()))
        bufs = [torch.empty(s, dtype=t.dtype, device=dev) for s in sizes]
        dist.all_gather(bufs, t.contiguous())
        return bufs

    pairs_all = gather_var(out.pairs)
    sims_all = gather_var(out.sims)
    idx_all = gather_var(out.nbr_idx.reshape(-1))
    val_all = gather_var(out.nbr_val.reshape(-1))
    # per-user prediction checksums (order-free inside a user): sum(col * 131 + pred) and count
    nt = out.pred_row_start.shape[1]
    cs = torch.zeros(n, dtype=torch.int64, device=dev)
    user_of = torch.repeat_interleave(torch.arange(n * nt, device=dev) // nt, out.pred_row_cnt.reshape(-1).long())
    starts = torch.repeat_interleave(out.pred_row_start.reshape(-1), out.pred_row_cnt.reshape(-1).long())
    within = torch.arange(user_of.numel(), device=dev) - torch.repeat_interleave(
        torch.cumsum(out.pred_row_cnt.reshape(-1).long(), 0) - out.pred_row_cnt.reshape(-1).long(),
        out.pred_row_cnt.reshape(-1).long())
    pos = starts + within
    cs.index_add_(0, user_of, out.pred_cols[pos].long() * 131 + out.pred_vals[pos].long())
    cs_all = gather_var(cs)
    ok = True
    msg = {}
    if rank == 0:
        rp, ci, va = synth.planted_csr(n * world, m, 0.05, seed=77, device=dev)
        single = LshCfPipeline(m=m, K=K, b=b, k=k, perm_tab=tab, device=dev, distributed=False).run(rp, ci, va)
        p1, order1 = torch.sort(single.pairs)
        pm, orderm = torch.sort(torch.cat(pairs_all))
        ok &= bool(torch.equal(p1, pm))
        ok &= bool(torch.equal(single.sims[order1], torch.cat(sims_all)[orderm]))
        ok &= bool(torch.equal(single.nbr_idx.reshape(-1), torch.cat(idx_all)))
        ok &= bool(torch.equal(single.nbr_val.reshape(-1), torch.cat(val_all)))
        nt1 = single.pred_row_start.shape[1]
        cnt1 = single.pred_row_cnt.reshape(-1).long()
        cs1 = torch.zeros(n * world, dtype=torch.int64, device=dev)
        user1 = torch.repeat_interleave(torch.arange(n * world * nt1, device=dev) // nt1, cnt1)
        st1 = torch.repeat_interleave(single.pred_row_start.reshape(-1), cnt1)
        w1 = torch.arange(user1.numel(), device=dev) - torch.repeat_interleave(torch.cumsum(cnt1, 0) - cnt1, cnt1)
        cs1.index_add_(0, user1, single.pred_cols[st1 + w1].long() * 131 + single.pred_vals[st1 + w1].long())
        ok &= bool(torch.equal(cs1, torch.cat(cs_all)))
        msg = {"world": world, "sets": n * world, "pairs": int(pm.numel()), "pairs_per_rank": [int(x.numel()) for x in pairs_all],
               "predictions_single": int(single.n_predictions), "identical": bool(ok)}
        print(json.dumps(msg))
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
