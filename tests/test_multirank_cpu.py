"""world_size-2 (gloo, CPU) test of the multi-rank plumbing of LshCfPipeline: row shards, signature
all-gather, bucket-record all-to-all keyed by the hash bits, scoring where the bucket lands, routing
of scored pairs to the owners of both ends, ratings all-gather.  The stage kernels are replaced by the
oracle-backed test double (tests/cpu_stage_double.py); the invariant is that candidate
set, rounded similarities, neighbour tables and predictions are identical for G = 1 and G = 2."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)

N_PER_RANK, M, K, B, TOPK = 80, 400, 32, 8, 7      # 80 % 32 != 0: a taste group straddles the two ranks


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run_pipeline(rank, world, rowptr, colidx, val, tab):
    import cpu_stage_double as double
    from query_recommendation_system_b200.pipeline import LshCfPipeline
    pipe = LshCfPipeline(m=M, K=K, b=B, k=TOPK, perm_tab=tab, device="cpu", stages=double)
    first = pipe.run(rowptr, colidx, val)
    learnt = dict(pipe._cap)
    again = pipe.run(rowptr, colidx, val)                  # second step: the capacities the first one learnt
    pipe._cap = {key: 8 for key in learnt}                 # far too small: every buffer overflows, the step is repeated
    forced = pipe.run(rowptr, colidx, val)
    for other in (again, forced):
        assert torch.equal(torch.sort(first.pairs).values, torch.sort(other.pairs).values)
        assert torch.equal(first.nbr_idx, other.nbr_idx) and torch.equal(first.nbr_val, other.nbr_val)
        assert first.n_predictions == other.n_predictions
    assert pipe._cap == learnt, "the repeated step must end with the same learnt capacities"
    return first


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import spec
        from query_recommendation_system_b200 import synth
        tab = spec.permutation_table(M, K, seed=7)
        rowptr, colidx, val = synth.planted_csr(N_PER_RANK, M, 0.06, seed=5, row_base=rank * N_PER_RANK)
        out = _run_pipeline(rank, world, rowptr, colidx, val, tab)
        np.savez(os.path.join(out_dir, "rank%d.npz" % rank),
                 pairs=out.pairs.numpy(), sims=out.sims.numpy(), idx=out.nbr_idx.numpy(), val=out.nbr_val.numpy(),
                 cnt=out.nbr_cnt.numpy(), rs=out.pred_row_start.numpy(), rc=out.pred_row_cnt.numpy(),
                 cols=out.pred_cols.numpy(), preds=out.pred_vals.numpy())
    finally:
        dist.destroy_process_group()


def _user_predictions(d, t):
    rs, rc = d["rs"], d["rc"]
    c = np.concatenate([d["cols"][rs[t, x]:rs[t, x] + rc[t, x]] for x in range(rs.shape[1])])
    p = np.concatenate([d["preds"][rs[t, x]:rs[t, x] + rc[t, x]] for x in range(rs.shape[1])])
    return c, p


@pytest.mark.timeout(600)
def test_two_ranks_equal_one_rank(tmp_path):
    from oracle import spec
    from query_recommendation_system_b200 import synth
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    ranks = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % r)) for r in range(world)]

    # single-rank truth from the oracle over the whole matrix
    n = N_PER_RANK * world
    rowptr, colidx, val = [x.numpy() for x in synth.planted_csr(n, M, 0.06, seed=5)]
    tab = spec.permutation_table(M, K, seed=7)
    sig = spec.signatures(rowptr, colidx, tab)
    cand = spec.candidates_packed(sig, B)
    assert cand.size > 50
    pr = np.stack([(cand >> np.uint64(32)).astype(np.int64), (cand & np.uint64(0xFFFFFFFF)).astype(np.int64)], 1)
    sims = spec.cosine_milli(sig, pr)

    # every candidate pair was scored on exactly one rank (where its first matching bucket landed)
    got = np.concatenate([r["pairs"].view(np.uint64) for r in ranks])
    got_s = np.concatenate([r["sims"] for r in ranks])
    order = np.argsort(got)
    assert np.array_equal(got[order], cand) and np.array_equal(got_s[order], sims)
    assert all(r["pairs"].size > 0 for r in ranks), "the bucket exchange left a rank without work"
    cross = ((pr[:, 0] // N_PER_RANK) != (pr[:, 1] // N_PER_RANK)).sum()
    assert cross > 0, "no candidate pair spans the two ranks: remote-row fetch would be untested"

    idx, v, cnt = spec.topk_tables(n, pr, sims / 1000.0, TOPK)
    ref = spec.predict_user_branch_csr(rowptr, colidx, val, M, idx, v, cnt)
    for r, d in enumerate(ranks):
        sl = slice(r * N_PER_RANK, (r + 1) * N_PER_RANK)
        assert np.array_equal(d["idx"], idx[sl]) and np.array_equal(d["val"], v[sl]) and np.array_equal(d["cnt"], cnt[sl])
        for t in range(N_PER_RANK):
            c, p = _user_predictions(d, t)
            assert np.array_equal(c, ref[r * N_PER_RANK + t][0]) and np.array_equal(p.astype(np.int64), ref[r * N_PER_RANK + t][1])


def test_single_rank_double_matches_oracle():
    """The same pipeline object with world size 1 (no process group) on the double — guards the G == 1 branch."""
    from oracle import spec
    from query_recommendation_system_b200 import synth
    rowptr, colidx, val = synth.planted_csr(N_PER_RANK, M, 0.06, seed=5)
    tab = spec.permutation_table(M, K, seed=7)
    out = _run_pipeline(0, 1, rowptr, colidx, val, tab)
    sig = spec.signatures(rowptr.numpy(), colidx.numpy(), tab)
    assert np.array_equal(out.sig.numpy().astype(np.int64), sig)
    assert np.array_equal(np.sort(out.pairs.numpy().view(np.uint64)), spec.candidates_packed(sig, B))
