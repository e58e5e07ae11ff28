"""GPU parity of the sparse predictor (recommender.py:36-47, 311-331) at the shapes bench.py runs and on the
paths the small cases never reach: the shipped column window with Q = 10 000, neighbour sub-segments of hundreds to
thousands of entries (many steps, several step-list rounds), odd nnz with exactly sized unaligned buffers, every
column format."""
import numpy as np
import pytest
import torch

from oracle import spec

pytestmark = pytest.mark.gpu


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda().contiguous()


def random_csr(n, m, lo, hi, seed, zero_frac=0.01):
    """Rows of lo..hi distinct ascending columns, ratings 1..100 with a few stored zeros."""
    rng = np.random.default_rng(seed)
    lens = rng.integers(lo, hi + 1, n)
    rowptr = np.zeros(n + 1, np.int64)
    np.cumsum(lens, out=rowptr[1:])
    colidx = np.empty(rowptr[-1], np.int32)
    for i in range(n):
        colidx[rowptr[i]:rowptr[i + 1]] = np.sort(rng.choice(m, lens[i], replace=False))
    val = rng.integers(1, 101, rowptr[-1]).astype(np.uint8)
    val[rng.random(val.size) < zero_frac] = 0
    return rowptr, colidx, val


def random_tables(n, k, seed, full=False, tie_heavy=False):
    rng = np.random.default_rng(seed)
    idx = np.full((n, k), -1, np.int32)
    sims = np.zeros((n, k))
    cnt = np.full(n, k, np.int32) if full else rng.integers(0, k + 1, n).astype(np.int32)
    for i in range(n):
        others = np.setdiff1d(np.arange(n), [i])
        idx[i, :cnt[i]] = rng.choice(others, cnt[i], replace=False)
        sims[i, :cnt[i]] = np.sort(rng.integers(0 if tie_heavy else 200, 1001, cnt[i]))[::-1] / 1000.0
    return idx, sims, cnt


def decode(rs, rc, cols, preds, cols_mode, tile, t):
    """(columns, predictions) of local user t from predict_sparse's outputs, whatever the column format."""
    nt = rs.shape[1]
    ps = np.concatenate([preds[rs[t, x]:rs[t, x] + rc[t, x]] for x in range(nt)])
    if cols_mode == 2:
        words = tile // 32
        bm = cols.view(np.uint32).reshape(rs.shape[0], nt, words)[t]
        bits = np.unpackbits(bm.view(np.uint8).reshape(nt, words * 4), axis=1, bitorder="little")
        cs = np.concatenate([x * tile + np.flatnonzero(bits[x]) for x in range(nt)])
        assert cs.size == ps.size
        return cs[ps != 0].astype(np.int64), ps[ps != 0].astype(np.int64)
    c = cols.view(np.uint16) if cols_mode == 1 else cols
    cs = np.concatenate([c[rs[t, x]:rs[t, x] + rc[t, x]] for x in range(nt)]).astype(np.int64)
    return cs, ps.astype(np.int64)


def run_and_check(rowptr, colidx, val, m, idx, sims, cnt, tile, integer, cols_mode=0, users=None, exact_alloc=False):
    from query_recommendation_system_b200 import stages
    n = rowptr.size - 1
    users = range(n) if users is None else users
    lo, hi = users[0], users[-1] + 1
    ref = spec.predict_user_branch_csr(rowptr, colidx, val, m, idx, sims, cnt, rows=range(lo, hi))
    total = sum(v[0].size for v in ref.values())
    sl = slice(lo, hi)
    tab = (cu(idx[sl]), cu(sims[sl]), cu(cnt[sl]))
    milli = cu(np.rint(sims[sl] * 1000).astype(np.int32)) if integer else None
    if exact_alloc:
        # buffers of exactly nnz elements whose end coincides with the end of a cudaMalloc'ed block would need a
        # raw allocation; the closest torch offers is a view that ends at the last element of its storage
        col_d = torch.empty(colidx.size + 3, dtype=torch.int32, device="cuda")[3:]
        val_d = torch.empty(val.size + 7, dtype=torch.uint8, device="cuda")[7:]
        col_d.copy_(torch.from_numpy(colidx))
        val_d.copy_(torch.from_numpy(val))
    else:
        col_d, val_d = cu(colidx), cu(val)
    rs, rc, cols, preds, cursor, nz = stages.predict_sparse(cu(rowptr), col_d, val_d, m, tab, lo, hi - lo, tile,
                                                            capacity=total + 64, u_milli=milli, cols_mode=cols_mode)
    assert int(nz.item()) == total
    rs, rc, cols, preds = rs.cpu().numpy(), rc.cpu().numpy(), cols.cpu().numpy(), preds.cpu().numpy()
    for t in range(hi - lo):
        gc, gp = decode(rs, rc, cols, preds, cols_mode, tile, t)
        assert np.array_equal(gc, ref[lo + t][0]), "columns of user %d" % (lo + t)
        assert np.array_equal(gp, ref[lo + t][1]), "predictions of user %d" % (lo + t)
    return total


@pytest.mark.parametrize("integer", [False, True])
@pytest.mark.parametrize("cols_mode", [0, 1, 2])
def test_long_segments_in_one_tile(integer, cols_mode):
    """Rows of 2 000 - 5 000 entries, k = 28, in the widest window a warp can own: every neighbour sub-segment is
    60 - 160 steps long, so one (user, window) needs many rounds of the step list."""
    from query_recommendation_system_b200 import _lib
    n, m, k = 96, 10_000, 28
    rowptr, colidx, val = random_csr(n, m, 2000, 5000, seed=41)
    idx, sims, cnt = random_tables(n, k, seed=42, full=True)
    tile = min(10_016, _lib.load().qrs_predict_max_tile(1 if integer else 0))
    total = run_and_check(rowptr, colidx, val, m, idx, sims, cnt, tile, integer, cols_mode, users=range(8, 72))
    assert total > 64 * 5000


@pytest.mark.parametrize("integer", [False, True])
def test_mixed_short_and_long_segments(integer):
    """Short (0 - 40) and long (1 500 - 3 000) rows mixed among the neighbours of one user, several tiles."""
    n, m, k = 128, 6000, 28
    rng = np.random.default_rng(5)
    a = random_csr(n, m, 0, 40, seed=43)
    b = random_csr(n, m, 1500, 3000, seed=44)
    pick = rng.random(n) < 0.5
    rows = [(a if pick[i] else b) for i in range(n)]
    lens = np.array([r[0][i + 1] - r[0][i] for i, r in enumerate(rows)])
    rowptr = np.zeros(n + 1, np.int64)
    np.cumsum(lens, out=rowptr[1:])
    colidx = np.concatenate([r[1][r[0][i]:r[0][i + 1]] for i, r in enumerate(rows)])
    val = np.concatenate([r[2][r[0][i]:r[0][i + 1]] for i, r in enumerate(rows)])
    idx, sims, cnt = random_tables(n, k, seed=45)
    for tile in (2048, 6016 if integer else 3008):
        run_and_check(rowptr, colidx, val, m, idx, sims, cnt, tile, integer)


@pytest.mark.parametrize("integer,tie_heavy", [(True, False), (True, True), (False, False)])
def test_shipped_tile_at_q10000(integer, tie_heavy):
    """The bench shape of the predictor: Q = 10 000 in the window choose_predict_tile ships (1 024 columns for
    the integer flavour: ten windows per user), rows of ~500 entries, k = 28."""
    from query_recommendation_system_b200 import stages, synth
    n, m, k = 1536, 10_000, 28
    rowptr, colidx, val = [x.numpy() for x in synth.planted_csr(n, m, 0.05, seed=20260102)]
    val = val.copy()
    if tie_heavy:
        rng = np.random.default_rng(3)
        val[:] = rng.choice(np.array([5, 15, 25, 35], np.uint8), size=val.size, p=[0.55, 0.25, 0.1, 0.1])
    idx = np.full((n, k), -1, np.int32)
    sims = np.zeros((n, k))
    rng = np.random.default_rng(9)
    cnt = rng.integers(k - 6, k + 1, n).astype(np.int32)
    for i in range(n):
        g0 = (i // 32) * 32
        cand = np.setdiff1d(np.arange(max(0, g0 - 16), min(n, g0 + 48)), [i])
        idx[i, :cnt[i]] = rng.choice(cand, cnt[i], replace=False)
        sims[i, :cnt[i]] = np.sort(rng.integers(0 if tie_heavy else 300, 1001, cnt[i]))[::-1] / 1000.0
    tile = stages.choose_predict_tile(m, colidx.size / n, integer)
    if integer:
        assert tile == 1024
    run_and_check(rowptr, colidx, val, m, idx, sims, cnt, tile, integer, cols_mode=2 if integer else 0,
                  users=range(512, 768))


@pytest.mark.parametrize("integer", [False, True])
def test_odd_nnz_exactly_sized_buffers(integer):
    """The last row ends at an odd position == nnz and the CSR buffers end right there (the paired
    8-byte column / 2-byte rating loads must not read past nnz nor need more than element alignment)."""
    n, m, k = 64, 700, 12
    rowptr, colidx, val = random_csr(n, m, 1, 61, seed=77)
    if rowptr[-1] % 2 == 0:                      # make nnz odd: drop the last entry
        rowptr[-1] -= 1
        colidx, val = colidx[:-1], val[:-1]
    assert rowptr[-1] % 2 == 1
    idx, sims, cnt = random_tables(n - 1, k, seed=78, full=True)          # neighbours among rows 0 .. n-2 ...
    idx = np.vstack([idx, idx[:1]])
    sims = np.vstack([sims, sims[:1]])
    cnt = np.append(cnt, cnt[0])
    idx[:n - 1, 0] = n - 1                                                 # ... plus the last row for everybody
    run_and_check(rowptr, colidx, val, m, idx, sims, cnt, 704, integer, exact_alloc=True)


@pytest.mark.parametrize("pack", ["class", "column"])
@pytest.mark.parametrize("integer", [False, True])
def test_prepared_row_layouts(pack, integer, monkeypatch):
    """Both layouts of the prepared rows (bank-class cells / column order, QRS_PREDICT_PACK) on columns that are all
    multiples of 16 — ONE bank class, so nearly every entry of a sub-segment is an 'extra' parked in another class's
    cell — with tie-heavy ratings, so that the half-way fix-up looks entries up in those sub-segments too; windows
    of 0 .. ~40 entries and a few of more than 128 (which keep the column order in either layout)."""
    monkeypatch.setenv("QRS_PREDICT_PACK", pack)
    n, m, k = 160, 16_384, 20
    rng = np.random.default_rng(123)
    lens = rng.integers(0, 400, n)
    lens[::17] = 1000                                  # a few rows with ~62 entries in every window ...
    rowptr = np.zeros(n + 1, np.int64)
    np.cumsum(lens, out=rowptr[1:])
    colidx = np.empty(rowptr[-1], np.int32)
    for i in range(n):
        colidx[rowptr[i]:rowptr[i + 1]] = 16 * np.sort(rng.choice(m // 16, lens[i], replace=False))
    big = np.sort(rng.choice(1024, 300, replace=False)).astype(np.int32)       # ... and one with 300 in window 0
    colidx = np.concatenate([colidx[:rowptr[1]], big, colidx[rowptr[2]:]])
    lens[1] = big.size
    np.cumsum(lens, out=rowptr[1:])
    val = rng.choice(np.array([5, 15, 25, 35], np.uint8), size=colidx.size, p=[0.55, 0.25, 0.1, 0.1])
    idx, sims, cnt = random_tables(n, k, seed=124, tie_heavy=True)
    idx[:, 0] = np.where(np.arange(n) == 1, 2, 1)      # the long row is everybody's first neighbour
    cnt = np.maximum(cnt, 1)
    sims[:, 0] = np.maximum(sims[:, 0], 0.5)
    run_and_check(rowptr, colidx, val, m, idx, sims, cnt, 1024, integer, cols_mode=1 if integer else 0)


class _TwoShardsOneDevice:
    """Stands in for stages.PeerWorkspace: the prepared-row buffers of two 'ranks' live on ONE device, so the
    sharded addressing of the predictor (global neighbour ids, rows read from another rank's buffer, staged or in
    place) is exercised without a second GPU."""

    def __init__(self, bufs, rank):
        import ctypes
        self.world, self.rank = len(bufs), rank
        self.local, self.nbytes = bufs[rank].data_ptr(), bufs[rank].numel()
        self.table = (ctypes.c_uint64 * self.world)(*[b.data_ptr() for b in bufs])


@pytest.mark.parametrize("integer,tie_heavy", [(True, False), (True, True), (False, False)])
@pytest.mark.parametrize("staging", ["in place", "staged", "staged: buffer too small"])
def test_rows_of_another_rank(integer, tie_heavy, staging):
    """Users of shard 0 whose neighbours live on both shards: rows of shard 1 read in place, copied into the
    staging buffer first (qrs_peer_rows_stage), or — buffer too small — announced and read in place after all."""
    from query_recommendation_system_b200 import stages
    n, m, k, tile = 512, 3000, 20, 1024
    h = n // 2
    rowptr, colidx, val = random_csr(n, m, 0, 300, seed=61)
    if tie_heavy:
        val = np.random.default_rng(62).choice(np.array([5, 15, 25, 35], np.uint8), size=val.size)
    idx, sims, cnt = random_tables(n, k, seed=63, tie_heavy=tie_heavy)
    shards = []
    for r in range(2):
        lo, hi = rowptr[r * h], rowptr[(r + 1) * h]
        shards.append((rowptr[r * h:(r + 1) * h + 1] - lo, colidx[lo:hi], val[lo:hi]))
    need = max(stages.predict_workspace_bytes(h, s[1].size, m, tile, 1 << 20) for s in shards)
    bufs = [torch.empty(need, dtype=torch.uint8, device="cuda") for _ in range(2)]
    plans = []
    for r, (rp, ci, va) in enumerate(shards):
        plans.append(stages.PredictPlan(cu(rp), cu(ci), cu(va), m, tile, capacity_hint=1 << 20,
                                        peers=_TwoShardsOneDevice(bufs, r)))
    ref = spec.predict_user_branch_csr(rowptr, colidx, val, m, idx, sims, cnt, rows=range(h))
    total = sum(v[0].size for v in ref.values())
    tab = (cu(idx[:h]), cu(sims[:h]), cu(cnt[:h]))
    milli = cu(np.rint(sims[:h] * 1000).astype(np.int32)) if integer else None
    stage = None
    if staging != "in place":
        probe = stages.PeerStage(plans[0], tab[0], tab[2], 0)                    # cap 0: only counts
        needed = int(probe.needed.item())
        remote_rows = idx[:h][(np.arange(k)[None, :] < cnt[:h, None]) & (idx[:h] >= h)] - h
        nt = (m + tile - 1) // tile
        steps_of = lambda r: sum(-(-np.count_nonzero((shards[1][1][shards[1][0][r]:shards[1][0][r + 1]] // tile) == t) // 32)
                                 for t in range(nt))
        assert needed == sum(steps_of(r) for r in remote_rows)
        stage = stages.PeerStage(plans[0], tab[0], tab[2], needed if staging == "staged" else needed // 2)
        assert bool(stage.buf[8:16].view(torch.int64).item()) == (staging == "staged")
    rp, ci, va = (cu(x) for x in shards[0])
    rs, rc, cols, preds, cursor, nz = stages.predict_sparse(rp, ci, va, m, tab, 0, h, tile, capacity=total + 64,
                                                            u_milli=milli, plan=plans[0], stage=stage)
    assert int(nz.item()) == total
    rs, rc, cols, preds = rs.cpu().numpy(), rc.cpu().numpy(), cols.cpu().numpy(), preds.cpu().numpy()
    for t in range(h):
        gc, gp = decode(rs, rc, cols, preds, 0, tile, t)
        assert np.array_equal(gc, ref[t][0]) and np.array_equal(gp, ref[t][1]), "user %d" % t
