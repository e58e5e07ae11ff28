"""GPU parity tests: every CUDA stage, called through the C ABI (ctypes → libqrs_b200.so),
against the CPU oracle (oracle/spec.py) and the golden vectors produced by the unmodified
reference (tests/golden/*.npz).  Bit-exact for signatures, bucket ids, candidate sets,
rounded similarities, neighbour tables and integer predictions."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import spec

pytestmark = pytest.mark.gpu

DEV = "cuda"


def cu(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.to(DEV).contiguous()


def unpack(pairs_t):
    p = pairs_t.cpu().numpy().view(np.uint64)
    return np.stack([(p >> np.uint64(32)).astype(np.int64), (p & np.uint64(0xFFFFFFFF)).astype(np.int64)], axis=1)


def gpu_signatures(g_rowptr, g_colidx, tab, b=0, mode=0, R=None):
    from query_recommendation_system_b200 import stages
    m, K = tab.shape
    n = g_rowptr.size - 1
    avg = g_colidx.size / max(n, 1)
    R = stages.choose_head_ranks(m, avg) if R is None else R
    tab_d = cu(tab, torch.int32)
    head = stages.perm_head(tab_d, R) if R > 0 else None
    return stages.minhash(cu(g_rowptr, torch.int64), cu(g_colidx, torch.int32), m, tab_d, head, b=b, mode=mode)


# ------------------------------------------------------------------------------ a1
TILE_MAX_ELEMENTS = 220000   # QRS_MINHASH_TILE: 32-bit words up to 32768 elements, then 16-bit words, then bytes


@pytest.mark.parametrize("mode", [0, 1, 2, 3])
def test_minhash_c1_bit_exact(golden_c1, mode):
    g = golden_c1
    tab = g["perm_tab"].astype(np.int32)
    sig, norm2, _ = gpu_signatures(g["rowptr"], g["colidx"], tab, mode=mode)
    ref = g["sig"].astype(np.int64)
    assert np.array_equal(sig.cpu().numpy().astype(np.int64), ref)
    assert np.array_equal(norm2.cpu().numpy(), (ref * ref).sum(axis=1))


@pytest.mark.parametrize("mode", [0, 1, 2, 3])
def test_minhash_synth_bit_exact(golden_synth, mode):
    g = golden_synth
    if mode == 3 and int(g["m"]) > TILE_MAX_ELEMENTS:
        pytest.skip("the tile scan's word map does not fit shared memory (covered by test_minhash_tile_refuses_large_m)")
    tab = spec.permutation_table(int(g["m"]), int(g["K"]), seed=int(g["perm_seed"]))
    sig, norm2, _ = gpu_signatures(g["rowptr"], g["colidx"], tab, mode=mode)
    ref = g["sig"].astype(np.int64)
    assert np.array_equal(sig.cpu().numpy().astype(np.int64), ref)    # includes empty sets → -1
    assert np.array_equal(norm2.cpu().numpy(), (ref * ref).sum(axis=1))


def test_minhash_head_exhaustion_falls_back_to_direct():
    # tiny head (R = 32) with short sets: most hashes are NOT resolved inside the head
    rng = np.random.default_rng(5)
    n, m, K = 300, 5000, 64
    rows = [np.sort(rng.choice(m, size=rng.integers(0, 40), replace=False)).astype(np.int32) for _ in range(n)]
    rowptr = np.zeros(n + 1, np.int64)
    rowptr[1:] = np.cumsum([r.size for r in rows])
    colidx = np.concatenate(rows).astype(np.int32)
    tab = spec.permutation_table(m, K, seed=3)
    ref = spec.signatures(rowptr, colidx, tab)
    for mode in (0, 1, 2, 3):
        sig, _, _ = gpu_signatures(rowptr, colidx, tab, mode=mode, R=32)
        assert np.array_equal(sig.cpu().numpy().astype(np.int64), ref)


def _ragged_rows(rng, n, m, lens):
    rows = [np.sort(rng.choice(m, size=int(l), replace=False)).astype(np.int32) for l in lens]
    rowptr = np.zeros(n + 1, np.int64)
    rowptr[1:] = np.cumsum([r.size for r in rows])
    colidx = np.concatenate(rows).astype(np.int32) if rowptr[-1] else np.zeros(0, np.int32)
    return rowptr, colidx


@pytest.mark.parametrize("mode", [0, 3])
@pytest.mark.parametrize("n,m,K,b", [(1, 64, 32, 8), (31, 900, 96, 24), (33, 900, 100, 25), (517, 3000, 180, 60),
                                       (1000, 10000, 128, 32), (200, 32768, 64, 16),
                                       (150, 32769, 64, 16), (90, 70000, 100, 20), (45, 104000, 256, 64),
                                       (70, 120000, 96, 32), (37, 200000, 40, 8)])
def test_minhash_tile_ragged_tiles(mode, n, m, K, b):
    """Tile scan (32 sets per CTA step up to 32768 elements, 16 sets with 16-bit words while they fit shared memory —
    values beyond 16 bits at m = 70000 / 104000 —, 8 sets with bytes beyond: m = 120000 / 200000): partial last tile,
    K not a multiple of 32, empty / one-element / short / long sets side by side in one tile (AUTO sends the short
    ones to the direct path), rows longer than one 256-entry build step per warp of a group, and the fused norm2 +
    bucket records."""
    from query_recommendation_system_b200 import stages
    rng = np.random.default_rng(n * 7 + K)
    kinds = rng.integers(0, 5, size=n)
    hi = max(2, min(m // 8, 600 if m <= 32768 else 2600))
    lens = np.where(kinds == 0, 0, np.where(kinds == 1, 1, np.where(kinds == 2, rng.integers(2, 12, size=n),
                                                                   rng.integers(hi // 2, hi, size=n))))
    if lens.sum() == 0:
        lens[0] = 3          # an all-empty CSR has no colidx buffer to pass
    rowptr, colidx = _ragged_rows(rng, n, m, lens)
    tab = spec.permutation_table(m, K, seed=11)
    ref = spec.signatures(rowptr, colidx, tab)
    sig, norm2, rec = gpu_signatures(rowptr, colidx, tab, b=b, mode=mode)
    assert np.array_equal(sig.cpu().numpy().astype(np.int64), ref)
    assert np.array_equal(norm2.cpu().numpy(), (ref * ref).sum(axis=1))
    rec2, norm2b = stages.band_records(sig, b, 0)
    assert torch.equal(rec, rec2) and torch.equal(norm2, norm2b)


@pytest.mark.parametrize("shift", [0, 1, 2, 3])
@pytest.mark.parametrize("m,K", [(3000, 64), (70000, 64)])
def test_minhash_tile_colidx_alignment(shift, m, K):
    """The tile scan's map build for every row misalignment (rows of all lengths mod 4), the array's last 16 bytes
    (exactly-sized tail) and a colidx pointer that is itself only 4-byte aligned gives the oracle's signatures —
    in the shipped 4-byte form and (test_minhash_tile_v4_csr_form re-runs this test with QRS_MINHASH_V4_CSR=1) in
    the 16-byte form, which walks each row from its start rounded down to 16 bytes."""
    from query_recommendation_system_b200 import stages
    rng = np.random.default_rng(1000 + shift + m)
    n = 133
    lens = rng.integers(250, 700, size=n)
    lens[:8] = [256, 257, 258, 259, 511, 512, 513, 1]
    rowptr, colidx = _ragged_rows(rng, n, m, lens)
    tab = spec.permutation_table(m, K, seed=5)
    ref = spec.signatures(rowptr, colidx, tab)
    tab_d = cu(tab, torch.int32)
    head = stages.perm_head(tab_d, stages.choose_head_ranks(m, colidx.size / n))
    buf = torch.empty(colidx.size + shift, dtype=torch.int32, device=DEV)
    buf[shift:] = torch.from_numpy(colidx).to(DEV)
    sig, norm2, _ = stages.minhash(cu(rowptr, torch.int64), buf[shift:], m, tab_d, head, b=0, mode=3)
    assert np.array_equal(sig.cpu().numpy().astype(np.int64), ref)
    assert np.array_equal(norm2.cpu().numpy(), (ref * ref).sum(axis=1))


def test_minhash_tile_v4_csr_form():
    """The kept 16-byte CSR-load form of the map build (QRS_MINHASH_V4_CSR=1, read once per process) is bit-exact too."""
    import os
    import subprocess
    import sys
    if os.environ.get("QRS_MINHASH_V4_CSR") == "1":
        pytest.skip("already inside the re-run")
    env = dict(os.environ, QRS_MINHASH_V4_CSR="1")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-m", "gpu", os.path.abspath(__file__), "-k",
                        "test_minhash_tile_colidx_alignment or test_minhash_tile_ragged_tiles"],
                       env=env, cwd=root, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-2000:]


def test_minhash_auto_beyond_the_tile_scan_splits_by_row_length():
    """m = 300 000 (no tile scan; the map of one set is 37 KB): AUTO runs the short and the empty sets in a map-less
    direct launch and the long ones (|S|^2 >= 16 m) in the per-set scan — every row written once, all bit-exact,
    records and norm2 included; DIRECT and FIRST_TOUCH on the same input agree."""
    from query_recommendation_system_b200 import stages
    rng = np.random.default_rng(77)
    m, K, b = 300_000, 32, 8
    lens = np.array([0, 1, 50, 800, 2000, 2190, 2191, 2500, 3000, 0, 40, 2600] * 5)
    rowptr, colidx = _ragged_rows(rng, lens.size, m, lens)
    tab = spec.permutation_table(m, K, seed=9)
    ref = spec.signatures(rowptr, colidx, tab)
    outs = []
    for mode in (0, 1, 2):
        sig, norm2, rec = gpu_signatures(rowptr, colidx, tab, b=b, mode=mode)
        assert np.array_equal(sig.cpu().numpy().astype(np.int64), ref), mode
        assert np.array_equal(norm2.cpu().numpy(), (ref * ref).sum(axis=1))
        outs.append(rec)
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])


def test_minhash_tile_equals_first_touch_at_bench_density():
    """C2-like rows (5 % of 10 000 elements, planted groups): the three scans agree, records included."""
    from query_recommendation_system_b200 import synth
    n, m, K, b = 4096 + 17, 10000, 128, 32
    rowptr, colidx, _ = synth.planted_csr(n, m, 0.05, seed=5)
    rowptr, colidx = rowptr.numpy(), colidx.numpy()
    tab = spec.permutation_table(m, K, seed=7)
    outs = [gpu_signatures(rowptr, colidx, tab, b=b, mode=mode) for mode in (2, 3, 0)]
    for sig, norm2, rec in outs[1:]:
        assert torch.equal(sig, outs[0][0]) and torch.equal(norm2, outs[0][1]) and torch.equal(rec, outs[0][2])
    ref = spec.signatures(rowptr[:257], colidx[:rowptr[256]], tab)
    assert np.array_equal(outs[1][0][:256].cpu().numpy().astype(np.int64), ref)


def test_minhash_tile_refuses_large_m():
    from query_recommendation_system_b200 import _lib
    rng = np.random.default_rng(2)
    n, m, K = 40, 300000, 32        # not even one byte per element fits shared memory
    rowptr, colidx = _ragged_rows(rng, n, m, rng.integers(50, 400, size=n))
    tab = spec.permutation_table(m, K, seed=1)
    with pytest.raises(_lib.QrsError):
        gpu_signatures(rowptr, colidx, tab, mode=3)
    sig, _, _ = gpu_signatures(rowptr, colidx, tab, mode=0)     # AUTO falls back to the per-set scan
    assert np.array_equal(sig.cpu().numpy().astype(np.int64), spec.signatures(rowptr, colidx, tab))


def test_minhash_seeded_5000x2000_and_fused_records():
    from query_recommendation_system_b200 import stages, synth
    n, m, K, b = 5000, 2000, 128, 32
    rowptr, colidx, _ = synth.planted_csr(n, m, 0.05, seed=20260101)
    rowptr, colidx = rowptr.numpy(), colidx.numpy()
    tab = spec.permutation_table(m, K, seed=7)
    sig, norm2, rec = gpu_signatures(rowptr, colidx, tab, b=b)
    ref = spec.signatures(rowptr, colidx, tab)
    assert np.array_equal(sig.cpu().numpy().astype(np.int64), ref)
    # the fused epilogue's records equal the stand-alone banding kernel's
    rec2, norm2b = stages.band_records(sig, b, 0)
    assert torch.equal(rec, rec2) and torch.equal(norm2, norm2b)
    ids = (rec.cpu().numpy().view(np.uint64) & np.uint64(0xFFFFFFFF)).astype(np.int64)
    assert np.array_equal(ids, np.broadcast_to(np.arange(n), (b, n)))
    bands = ((rec.cpu().numpy().view(np.uint64) >> np.uint64(32)) & np.uint64(0xFF)).astype(np.int64)
    assert np.array_equal(bands, np.broadcast_to(np.arange(b)[:, None], (b, n)))


# --------------------------------------------------------------------------- a2/a3
def test_band_keys_exact(golden_synth):
    from query_recommendation_system_b200 import stages
    g = golden_synth
    sig = g["sig"].astype(np.int64)
    sig_d = cu(sig, torch.int32)
    for b in g["bands"].tolist():
        r = sig.shape[1] // b
        keys, hi, skipped = stages.band_keys(sig_d, b)
        assert np.array_equal(skipped.cpu().numpy().astype(bool), spec.band_is_skipped(sig, b))
        lanes = spec.band_lanes(sig, b).astype(np.uint16).astype(np.uint64)
        lo_ref = np.zeros(lanes.shape[:2], np.uint64)
        hi_ref = np.zeros(lanes.shape[:2], np.uint64)
        for c in range(r):
            hi_ref = (hi_ref << np.uint64(16)) | (lo_ref >> np.uint64(48))
            lo_ref = (lo_ref << np.uint64(16)) | lanes[:, :, c]
        assert np.array_equal(keys.cpu().numpy().view(np.uint64), lo_ref)
        if r <= 4:
            assert hi is None
            assert np.array_equal(lo_ref, spec.band_keys_u64(sig, b))
        else:
            assert np.array_equal(hi.cpu().numpy().view(np.uint64), hi_ref)


def test_band_keys_rejects_bad_band_count():
    from query_recommendation_system_b200 import stages
    with pytest.raises(AssertionError):                       # lsh.py:20
        stages.band_keys(torch.zeros((4, 10), dtype=torch.int32, device=DEV), 3)


# --------------------------------------------------------------------------- a3/a4
@pytest.mark.parametrize("n_per_seg,segs,bit_lo,bit_hi", [
    (1, 1, 0, 64), (37, 5, 40, 64), (4096, 3, 32, 64), (4097, 2, 40, 61), (100_000, 4, 40, 64), (250_000, 1, 0, 64)])
def test_sort_matches_stable_numpy(n_per_seg, segs, bit_lo, bit_hi):
    from query_recommendation_system_b200 import stages
    rng = np.random.default_rng(n_per_seg + segs)
    keys = rng.integers(0, 2 ** 63, size=(segs, n_per_seg), dtype=np.int64).view(np.uint64)
    keys ^= rng.integers(0, 2, size=keys.shape, dtype=np.uint64) << np.uint64(63)
    keys[:, ::3] &= np.uint64(0xFFFF0000FFFFFFFF)             # many equal digits → stability matters
    field = (keys >> np.uint64(bit_lo)) & np.uint64((1 << (bit_hi - bit_lo)) - 1 if bit_hi - bit_lo < 64 else 2 ** 64 - 1)
    ref = np.stack([keys[s][np.argsort(field[s], kind="stable")] for s in range(segs)])
    got = stages.sort_u64(cu(keys.view(np.int64)).reshape(-1), n_per_seg, segs, bit_lo, bit_hi)
    assert np.array_equal(got.cpu().numpy().view(np.uint64).reshape(segs, n_per_seg), ref)


# ------------------------------------------------------------------------ a5 + a8
def _gpu_candidates(sig_np, b, want_sims=True):
    from query_recommendation_system_b200.lsh import LSH
    l = LSH(b)
    l.compute_buckets_batch(cu(sig_np, torch.int32))
    pairs, sims = l.candidates_scored(want_sims)
    return pairs, sims


def test_candidates_and_cosine_c1(golden_c1):
    g = golden_c1
    sig = g["sig"].astype(np.int64)
    pairs, sims = _gpu_candidates(sig, int(g["b"]))
    got = unpack(pairs)
    order = np.lexsort((got[:, 1], got[:, 0]))
    assert np.array_equal(got[order], g["candidates"].astype(np.int64))         # 85 pairs, each once
    assert np.array_equal(sims.cpu().numpy()[order], spec.cosine_milli(sig, g["candidates"].astype(np.int64)))


def test_candidates_and_cosine_synth(golden_synth):
    g = golden_synth
    sig = g["sig"].astype(np.int64)
    for b in g["bands"].tolist():
        p = "b%d_" % b
        pairs, sims = _gpu_candidates(sig, b)
        got = unpack(pairs)
        order = np.lexsort((got[:, 1], got[:, 0]))
        ref = g[p + "candidates"].astype(np.int64)
        assert np.array_equal(got[order], ref), "candidate set differs for b=%d" % b
        assert np.array_equal(sims.cpu().numpy()[order] / 1000.0, g[p + "pair_sims"])


def test_candidates_seeded_planted_vs_oracle():
    from query_recommendation_system_b200 import synth
    n, m, K = 4096, 2000, 128
    rowptr, colidx, _ = synth.planted_csr(n, m, 0.05, seed=20260102)
    tab = spec.permutation_table(m, K, seed=7)
    sig = spec.signatures(rowptr.numpy(), colidx.numpy(), tab)
    for b in (16, 32, 64):
        pairs, sims = _gpu_candidates(sig, b)
        got = np.sort(pairs.cpu().numpy().view(np.uint64))
        ref = spec.candidates_packed(sig, b)
        assert np.array_equal(got, ref)
        srt = np.argsort(pairs.cpu().numpy().view(np.uint64))
        assert np.array_equal(sims.cpu().numpy()[srt], spec.cosine_milli(sig, unpack(pairs)[srt]))


def test_candidates_k256_b64_fast_path_vs_oracle():
    # r = 4 with K = 256: two 128-element chunks per row in the 16-byte-load kernel
    from query_recommendation_system_b200 import synth
    n, m, K, b = 2048, 3000, 256, 64
    rowptr, colidx, _ = synth.planted_csr(n, m, 0.04, seed=20260105)
    tab = spec.permutation_table(m, K, seed=7)
    sig = spec.signatures(rowptr.numpy(), colidx.numpy(), tab)
    sig[5] = -1                                  # an empty set: all bands all -1
    sig[40:44, 128:132] = 65535                  # a band whose lanes wrap to -1 (lsh.py:28, 47) in the second chunk
    pairs, sims = _gpu_candidates(sig, b)
    got = np.sort(pairs.cpu().numpy().view(np.uint64))
    assert np.array_equal(got, spec.candidates_packed(sig, b))
    srt = np.argsort(pairs.cpu().numpy().view(np.uint64))
    assert np.array_equal(sims.cpu().numpy()[srt], spec.cosine_milli(sig, unpack(pairs)[srt]))


def test_candidates_degenerate_inputs():
    # all sets empty → every band all -1 → no candidates (lsh.py:47); all identical → all pairs
    K, b = 32, 8
    empty = np.full((50, K), -1, np.int64)
    pairs, _ = _gpu_candidates(empty, b)
    assert pairs.numel() == 0
    same = np.tile(np.arange(K, dtype=np.int64) * 7 + 1, (40, 1))
    pairs, sims = _gpu_candidates(same, b)
    assert pairs.numel() == 40 * 39 // 2 and bool((sims == 1000).all())
    # values congruent to 65535 mod 65536 are swallowed by the -1 rule
    wrap = np.full((10, K), 65535, np.int64)
    pairs, _ = _gpu_candidates(wrap, b)
    assert pairs.numel() == 0 and len(spec.candidates(wrap, b)) == 0


def test_lsh_class_matches_reference_surface(golden_c1):
    from query_recommendation_system_b200 import LSH
    g = golden_c1
    sig = g["sig"].astype(np.int64)
    lsh = LSH(int(g["b"]))
    for s in sig:                                  # recommender.py:173-174
        lsh.compute_buckets(s)
    assert lsh.counter == sig.shape[0]
    cand = lsh.get_candidates(sig)
    assert cand == set(map(tuple, g["candidates"].tolist()))
    sub = lsh.make_subvecs(sig[3])
    assert sub.dtype == np.int16 and sub.shape == (int(g["b"]), sig.shape[1] // int(g["b"]))
    # bucket partition (str key → ids), as the reference's LSH.buckets holds it
    ref = {}
    for t in range(g["bucket_band"].size):
        key = ",".join(str(int(x)) for x in g["bucket_lanes"][t])
        ref[(int(g["bucket_band"][t]), key)] = g["bucket_members"][g["bucket_offs"][t]:g["bucket_offs"][t + 1]].tolist()
    got = {(band, k): v for band, d in enumerate(lsh.buckets) for k, v in d.items()}
    assert got == ref
    with pytest.raises(AssertionError):
        LSH(7).compute_buckets(sig[0])             # 180 % 7 != 0, lsh.py:20


# ------------------------------------------------------------------------ a7 + a9
def test_topk_tables_canonical(golden_synth):
    from query_recommendation_system_b200 import stages
    g = golden_synth
    n, k = int(g["n"]), int(g["k"])
    for b in g["bands"].tolist():
        p = "b%d_" % b
        pairs = g[p + "candidates"].astype(np.int64)
        sims = g[p + "pair_sims"]
        packed = (pairs[:, 0].astype(np.uint64) << np.uint64(32)) | pairs[:, 1].astype(np.uint64)
        milli = np.rint(sims * 1000).astype(np.int32)
        idx, val, cnt = stages.topk_tables(cu(packed.view(np.int64)), cu(milli), pairs.shape[0], 0, n, k)
        ridx, rval, rcnt = spec.topk_tables(n, pairs, sims, k)
        assert np.array_equal(cnt.cpu().numpy(), rcnt)
        assert np.array_equal(idx.cpu().numpy(), ridx)
        assert np.array_equal(val.cpu().numpy(), rval)          # bit-identical fp64 (milli / 1000.0)


@pytest.mark.parametrize("k", [1, 11, 28, 50, 64])
def test_topk_heavy_rows_and_row_ranges(k):
    from query_recommendation_system_b200 import stages
    rng = np.random.default_rng(k)
    n = 700
    # row 0 is adjacent to everything (degree 699 ≫ 64), plus random pairs with many tied sims
    a = rng.integers(0, n, 6000)
    c = rng.integers(0, n, 6000)
    keep = a != c
    lo, hi = np.minimum(a, c)[keep], np.maximum(a, c)[keep]
    star = np.stack([np.zeros(n - 1, np.int64), np.arange(1, n)], axis=1)
    pairs = np.unique(np.concatenate([np.stack([lo, hi], axis=1), star]), axis=0)
    sims = rng.integers(0, 40, pairs.shape[0]) / 1000.0 + 0.9
    packed = (pairs[:, 0].astype(np.uint64) << np.uint64(32)) | pairs[:, 1].astype(np.uint64)
    milli = np.rint(sims * 1000).astype(np.int32)
    ridx, rval, rcnt = spec.topk_tables(n, pairs, milli / 1000.0, k)
    for row_base, n_rows in ((0, n), (100, 250), (650, 50)):
        idx, val, cnt = stages.topk_tables(cu(packed.view(np.int64)), cu(milli), pairs.shape[0], row_base, n_rows, k)
        assert np.array_equal(cnt.cpu().numpy(), rcnt[row_base:row_base + n_rows])
        assert np.array_equal(idx.cpu().numpy(), ridx[row_base:row_base + n_rows])
        assert np.array_equal(val.cpu().numpy(), rval[row_base:row_base + n_rows])


# ---------------------------------------------------------------------- a10 + a11
def _dense_to_csr(R):
    U, Q = R.shape
    rowptr = np.zeros(U + 1, np.int64)
    rowptr[1:] = np.cumsum((R != 0).sum(axis=1))
    rows, cols = np.nonzero(R)
    return rowptr, cols.astype(np.int32), R[rows, cols].astype(np.uint8)


def _table(idx, val, cnt):
    return cu(idx, torch.int32), cu(val, torch.float64), cu(np.maximum(cnt, 0), torch.int32)


def test_predict_dense_c1_matches_reference(golden_c1):
    from query_recommendation_system_b200 import stages
    g = golden_c1
    R = g["ratings"].astype(np.int64)
    rowptr, colidx, val = _dense_to_csr(R)
    out = stages.predict_dense(cu(rowptr), cu(colidx), cu(val), R.shape[1],
                               _table(g["q_idx"], g["q_val"], g["q_cnt"]),
                               _table(g["u_idx"], g["u_val"], g["u_cnt"]))
    got = out.cpu().numpy().astype(np.int64)
    assert np.array_equal(got, g["predictions"].astype(np.int64))           # all 7450 predicted cells + ratings
    assert np.array_equal(np.argwhere(got == 0), g["missed"])               # recommender.py:341
    # single-branch variants against the oracle
    for q_on, u_on in ((True, False), (False, True)):
        qt = (g["q_idx"], g["q_val"], np.maximum(g["q_cnt"], 0)) if q_on else None
        ut = (g["u_idx"], g["u_val"], np.maximum(g["u_cnt"], 0)) if u_on else None
        ref = spec.predict_dense(R, qt, ut)
        out = stages.predict_dense(cu(rowptr), cu(colidx), cu(val), R.shape[1],
                                   _table(*qt) if qt else None, _table(*ut) if ut else None)
        assert np.array_equal(out.cpu().numpy().astype(np.int64), ref)


def test_predict_dense_half_way_ties():
    # ratings all 5: u*0.7 + 18 = 21.5 up to rounding noise — the tie the golden wavg cases probe
    from query_recommendation_system_b200 import stages
    rng = np.random.default_rng(17)
    U, Q, k = 60, 90, 9
    R = (rng.random((U, Q)) < 0.4).astype(np.int64) * 5
    R[::7] = (rng.random((R[::7].shape)) < 0.4) * rng.integers(1, 101, R[::7].shape)
    idx = np.stack([rng.choice(U, k, replace=False) for _ in range(U)]).astype(np.int32)
    val = np.sort(np.around(rng.random((U, k)), 3), axis=1)[:, ::-1].copy()
    cnt = rng.integers(0, k + 1, U).astype(np.int32)
    qidx = np.stack([rng.choice(Q, k, replace=False) for _ in range(Q)]).astype(np.int32)
    qval = np.sort(np.around(rng.random((Q, k)), 3), axis=1)[:, ::-1].copy()
    qcnt = rng.integers(0, k + 1, Q).astype(np.int32)
    ref = spec.predict_dense(R, (qidx, qval, qcnt), (idx, val, cnt))
    rowptr, colidx, v = _dense_to_csr(R)
    out = stages.predict_dense(cu(rowptr), cu(colidx), cu(v), Q, _table(qidx, qval, qcnt), _table(idx, val, cnt))
    assert np.array_equal(out.cpu().numpy().astype(np.int64), ref)


def _sparse_case(n, m, k, seed, tie_heavy, pool=32):
    from query_recommendation_system_b200 import synth
    rowptr_t, colidx_t, val_t = synth.planted_csr(n, m, 0.05, seed=seed)
    rowptr, colidx, val = rowptr_t.numpy(), colidx_t.numpy(), val_t.numpy().copy()
    rng = np.random.default_rng(seed)
    if tie_heavy:
        # ratings from {5, 15, 25, 35}: 0.7 * u + 18 lands exactly on .5 whenever the raters agree
        val[:] = rng.choice(np.array([5, 15, 25, 35], np.uint8), size=val.size, p=[0.55, 0.25, 0.1, 0.1])
        val[rng.random(val.size) < 0.02] = 0          # a few stored zeros (ignored like unrated cells)
    idx = np.full((n, k), -1, np.int32)
    sims = np.zeros((n, k))
    cnt = rng.integers(0, k + 1, n).astype(np.int32)
    for i in range(n):
        g0 = (i // pool) * pool
        cand = np.setdiff1d(np.arange(g0, g0 + pool), [i])
        idx[i, :cnt[i]] = rng.choice(cand, cnt[i], replace=False)
        sims[i, :cnt[i]] = np.sort(rng.integers(0 if tie_heavy else 300, 1001, cnt[i]))[::-1] / 1000.0
    return rowptr, colidx, val, idx, sims, cnt


@pytest.mark.parametrize("tile,integer", [(256, False), (512, False), (1024, False), (256, True), (512, True), (1024, True)])
@pytest.mark.parametrize("tie_heavy", [False, True])
def test_predict_sparse_matches_oracle(tile, integer, tie_heavy):
    from query_recommendation_system_b200 import stages
    n, m, k = 1024, 3000, 19
    rowptr, colidx, val, idx, sims, cnt = _sparse_case(n, m, k, 20260103 + int(tie_heavy), tie_heavy)
    user_base, n_users = 256, 512
    ref = spec.predict_user_branch_csr(rowptr, colidx, val, m, idx, sims, cnt, rows=range(user_base, user_base + n_users))
    total = sum(v[0].size for v in ref.values())
    sl = slice(user_base, user_base + n_users)
    tab = (cu(idx[sl]), cu(sims[sl]), cu(cnt[sl]))
    milli = cu(np.rint(sims[sl] * 1000).astype(np.int32)) if integer else None
    rs, rc, cols, preds, cursor, nz = stages.predict_sparse(cu(rowptr), cu(colidx), cu(val), m, tab, user_base, n_users,
                                                        tile, capacity=total + 5, u_milli=milli)
    assert int(cursor.item()) == total and int(nz.item()) == total
    rs, rc, cols, preds = rs.cpu().numpy(), rc.cpu().numpy(), cols.cpu().numpy(), preds.cpu().numpy()
    n_tie = 0
    for t in range(n_users):
        gc = np.concatenate([cols[rs[t, c]:rs[t, c] + rc[t, c]] for c in range(rs.shape[1])])
        gp = np.concatenate([preds[rs[t, c]:rs[t, c] + rc[t, c]] for c in range(rs.shape[1])])
        rcols, rpreds = ref[user_base + t]
        assert np.array_equal(gc, rcols) and np.array_equal(gp.astype(np.int64), rpreds)
    # counting pass (capacity 0) reports the same total without writing
    _, _, _, _, cursor0, _ = stages.predict_sparse(cu(rowptr), cu(colidx), cu(val), m, tab, user_base, n_users, tile, 0,
                                                u_milli=milli)
    assert int(cursor0.item()) == total


def test_topk_milli_table_is_exact(golden_synth):
    from query_recommendation_system_b200 import stages
    g = golden_synth
    n, k, b = int(g["n"]), int(g["k"]), int(g["bands"][0])
    pairs = g["b%d_candidates" % b].astype(np.int64)
    packed = (pairs[:, 0].astype(np.uint64) << np.uint64(32)) | pairs[:, 1].astype(np.uint64)
    milli = np.rint(g["b%d_pair_sims" % b] * 1000).astype(np.int32)
    idx, val, cnt, mil = stages.topk_tables(cu(packed.view(np.int64)), cu(milli), pairs.shape[0], 0, n, k, want_milli=True)
    assert np.array_equal(mil.cpu().numpy() / 1000.0, val.cpu().numpy())


def test_cpu_double_record_hash_matches_cuda():
    """The CPU test double used for the gloo multi-rank tests builds bucket records with the same
    hash24 | band | id layout as the CUDA kernels (csrc/common.cuh)."""
    import cpu_stage_double as double
    from query_recommendation_system_b200 import stages
    g = load_golden("synth_wrap")
    sig = g["sig"].astype(np.int64)
    for b in (4, 8, 16):
        rec, _ = stages.band_records(cu(sig, torch.int32), b, 1000)
        assert np.array_equal(rec.cpu().numpy().view(np.uint64), double.band_records_np(sig, b, 1000))


@pytest.mark.parametrize("n,K,b", [(1, 128, 32), (63, 128, 32), (64, 128, 32), (65, 128, 32), (1000, 256, 64), (777, 128, 64),
                                   (300, 96, 32), (130, 128, 16), (200, 64, 64), (90, 126, 42), (5000, 128, 32)])
def test_band_records_tile_form(n, K, b):
    """qrs_band_records in its tile form (signature tiles staged by bulk asynchronous copies, records transposed in
    shared memory): partial tiles, r = 1 / 2 / 3 / 4 / 8, -1 rows (sentinel hash), int16 wrap; K = 126 takes the
    per-row kernel (a tile is not a 16-byte multiple).  Against the NumPy record layout, norm2 against int64 sums."""
    import cpu_stage_double as double
    from query_recommendation_system_b200 import stages
    rng = np.random.default_rng(n * 131 + K + b)
    sig = rng.integers(0, 70000, size=(n, K)).astype(np.int64)
    sig[rng.random(n) < 0.1] = -1                      # empty sets
    if n > 2:
        sig[2, : K // b] = -1                          # one all-(-1) band inside a regular row
    rec, norm2 = stages.band_records(cu(sig, torch.int32), b, 12345)
    assert np.array_equal(rec.cpu().numpy().view(np.uint64), double.band_records_np(sig, b, 12345))
    assert np.array_equal(norm2.cpu().numpy(), (sig * sig).sum(axis=1))
    rec_only, none = stages.band_records(cu(sig, torch.int32), b, 0, want_norm2=False)
    assert none is None and np.array_equal(rec_only.cpu().numpy().view(np.uint64), double.band_records_np(sig, b, 0))


def test_predict_sparse_regions_and_overflow():
    """Slices reserved from several cursors: same per-user content; a too-small region drops writes but
    still reports the demand, so the caller can size the buffer and run again."""
    from query_recommendation_system_b200 import stages
    n, m, k = 512, 2000, 11
    rowptr, colidx, val, idx, sims, cnt = _sparse_case(n, m, k, 77, False)
    ref = spec.predict_user_branch_csr(rowptr, colidx, val, m, idx, sims, cnt)
    total = sum(v[0].size for v in ref.values())
    tab = (cu(idx), cu(sims), cu(cnt))
    milli = cu(np.rint(sims * 1000).astype(np.int32))
    regions = 8
    _, _, _, _, demand, _ = stages.predict_sparse(cu(rowptr), cu(colidx), cu(val), m, tab, 0, n, 512, 0,
                                               u_milli=milli, n_regions=regions)
    assert int(demand.sum().item()) == total
    rcap = int(demand.max().item())
    rs, rc, cols, preds, cur, nz = stages.predict_sparse(cu(rowptr), cu(colidx), cu(val), m, tab, 0, n, 512, rcap * regions,
                                                     u_milli=milli, n_regions=regions)
    assert torch.equal(cur, demand)
    rs, rc, cols, preds = rs.cpu().numpy(), rc.cpu().numpy(), cols.cpu().numpy(), preds.cpu().numpy()
    for t in range(n):
        gc = np.concatenate([cols[rs[t, c]:rs[t, c] + rc[t, c]] for c in range(rs.shape[1])])
        gp = np.concatenate([preds[rs[t, c]:rs[t, c] + rc[t, c]] for c in range(rs.shape[1])])
        assert np.array_equal(gc, ref[t][0]) and np.array_equal(gp.astype(np.int64), ref[t][1])
    # every slice lies inside its region
    reg = rs // rcap
    assert ((rs + rc - 1)[rc > 0] // rcap == reg[rc > 0]).all()


@pytest.mark.parametrize("N", [1, 10, 40, 64])
def test_recommend_topn_matches_numpy(N):
    """f-2: per-user top-N over the sparse predictions == argsort of the same user's predictions
    (prediction descending, column ascending), recommender.py:345-374."""
    from query_recommendation_system_b200 import stages
    n, m, k = 608, 2500, 15          # a multiple of the 32-row taste groups the neighbour ids are drawn from
    rowptr, colidx, val, idx, sims, cnt = _sparse_case(n, m, k, 91, False)
    ref = spec.predict_user_branch_csr(rowptr, colidx, val, m, idx, sims, cnt)
    total = sum(v[0].size for v in ref.values())
    tab = (cu(idx), cu(sims), cu(cnt))
    for tile, mode in ((512, 0), (1024, 1)):
        rs, rc, cols, preds, _, _ = stages.predict_sparse(cu(rowptr), cu(colidx), cu(val), m, tab, 0, n, tile, total + 8,
                                                          u_milli=cu(np.rint(sims * 1000).astype(np.int32)), cols_mode=mode)
        tidx, tval, tcnt = [x.cpu().numpy() for x in stages.recommend_topn(rs, rc, cols, preds, N)]
        for u in range(n):
            c, p = ref[u]
            order = np.lexsort((c, -p))[:N]
            assert tcnt[u] == order.size
            assert np.array_equal(tidx[u, :order.size], c[order]) and np.array_equal(tval[u, :order.size], p[order])
            assert (tidx[u, order.size:] == -1).all() and (tval[u, order.size:] == 0).all()


def test_answer_sets_match_numpy_masks():
    """f-1: conjunctive attr = value predicates over a dictionary-encoded table (recommender.py:76-95), incl.
    wildcard-only queries (every tuple), impossible values (empty answer) and a tuple count that is not a
    multiple of 32."""
    from query_recommendation_system_b200 import stages
    rng = np.random.default_rng(12)
    drows, F, nq = 4999, 5, 300
    card = [3, 40, 7, 90, 2]
    table = np.stack([rng.integers(0, c, drows) for c in card]).astype(np.int32)       # [F, drows]
    queries = np.full((nq, F), -1, np.int32)
    for q in range(nq):
        for f in rng.choice(F, size=rng.integers(0, 4), replace=False):
            queries[q, f] = rng.integers(0, card[f])
    queries[7, 2] = -2                                   # a value that does not occur
    rowptr, colidx = stages.answer_sets(cu(table), cu(queries))
    rowptr, colidx = rowptr.cpu().numpy(), colidx.cpu().numpy()
    for q in range(nq):
        mask = np.ones(drows, bool)
        for f in range(F):
            if queries[q, f] != -1:
                mask &= table[f] == queries[q, f]
        assert np.array_equal(colidx[rowptr[q]:rowptr[q + 1]], np.flatnonzero(mask))
    assert rowptr[8] == rowptr[7]                        # the impossible query has an empty answer set
    assert rowptr[-1] == colidx.size


def test_pair_cosine_explicit_list(golden_synth):
    """qrs_pair_cosine: rounded cosine of an explicit pair list == the oracle's np.around(cos, 3) * 1000."""
    from query_recommendation_system_b200 import stages
    g = golden_synth
    sig = g["sig"].astype(np.int64)
    b = int(g["bands"][0])
    pairs = g["b%d_candidates" % b].astype(np.int64)
    packed = (pairs[:, 0].astype(np.uint64) << np.uint64(32)) | pairs[:, 1].astype(np.uint64)
    sig_d = cu(sig, torch.int32)
    _, norm2 = stages.band_records(sig_d, b, 0)
    got = stages.pair_cosine(sig_d, norm2, cu(packed.view(np.int64)))
    assert np.array_equal(got.cpu().numpy() / 1000.0, g["b%d_pair_sims" % b])


@pytest.mark.parametrize("integer", [False, True])
def test_predict_sparse_wide_tables_k50(integer):
    """Neighbour tables wider than one warp (k = 50, the 10M-user config's top-k): two metadata rounds."""
    from query_recommendation_system_b200 import stages
    n, m, k = 512, 2000, 50
    rowptr, colidx, val, idx, sims, cnt = _sparse_case(n, m, k, 5150, False, pool=64)
    assert cnt.max() > 32
    ref = spec.predict_user_branch_csr(rowptr, colidx, val, m, idx, sims, cnt)
    total = sum(v[0].size for v in ref.values())
    milli = cu(np.rint(sims * 1000).astype(np.int32)) if integer else None
    rs, rc, cols, preds, cursor, nz = stages.predict_sparse(cu(rowptr), cu(colidx), cu(val), m, (cu(idx), cu(sims), cu(cnt)),
                                                            0, n, 704, total + 3, u_milli=milli)
    assert int(cursor.item()) == total
    rs, rc, cols, preds = rs.cpu().numpy(), rc.cpu().numpy(), cols.cpu().numpy(), preds.cpu().numpy()
    for t in range(n):
        gc = np.concatenate([cols[rs[t, c]:rs[t, c] + rc[t, c]] for c in range(rs.shape[1])])
        gp = np.concatenate([preds[rs[t, c]:rs[t, c] + rc[t, c]] for c in range(rs.shape[1])])
        assert np.array_equal(gc, ref[t][0]) and np.array_equal(gp.astype(np.int64), ref[t][1])


# ------------------------------------------------------------------------------ f-3
def _ratings_csr(n, m, fill, seed):
    """Random ratings rows incl. an empty row, a single-rating row, a constant row and explicit 0 cells."""
    rng = np.random.default_rng(seed)
    lens = rng.binomial(m, fill, n)
    lens[0], lens[1] = 0, 1
    rowptr = np.zeros(n + 1, np.int64)
    rowptr[1:] = np.cumsum(lens)
    colidx = np.concatenate([np.sort(rng.choice(m, l, replace=False)) for l in lens]).astype(np.int32)
    val = rng.integers(1, 101, colidx.size).astype(np.uint8)
    val[rowptr[2]:rowptr[3]] = 50                          # constant row: centres to zero norm
    val[rng.random(val.size) < 0.02] = 0                   # stored but "not rated" cells
    val[rowptr[4]:rowptr[5]] = np.where(np.arange(lens[4]) % 2 == 0, 1, 255)     # both ends of the uint8 range
    return rowptr, colidx, val


def test_center_rows_matches_oracle():
    from query_recommendation_system_b200 import stages
    rowptr, colidx, val = _ratings_csr(300, 900, 0.1, seed=2)
    cval, n2 = stages.center_rows(cu(rowptr), cu(val))
    want_c, want_n2 = spec.centered_rows(rowptr, val)
    assert np.array_equal(cval.cpu().numpy(), want_c) and np.array_equal(n2.cpu().numpy(), want_n2)
    assert want_n2[0] == 0 and want_n2[1] == 0 and want_n2[2] == 0 and want_c.min() < -100 and want_c.max() > 100


@pytest.mark.parametrize("m,fill", [(300, 0.2), (70_000, 0.002), (250_000, 0.0006)])
def test_pair_centered_cosine_matches_oracle(m, fill):
    """m = 300 / 70 000: row i staged in shared memory; m = 250 000: beyond the staged limit, search kernel."""
    from query_recommendation_system_b200 import stages
    n = 200
    rowptr, colidx, val = _ratings_csr(n, m, fill, seed=m)
    rng = np.random.default_rng(5)
    ii, jj = np.triu_indices(n, 1)
    pick = rng.choice(ii.size, 6000, replace=False)
    pr = np.stack([ii[pick], jj[pick]], 1)
    pr = pr[np.argsort(pr[:, 0], kind="stable")]           # grouped by first id, seconds in random order
    want = spec.centered_cosine_milli(rowptr, colidx, val, pr)
    packed = cu(((pr[:, 0].astype(np.uint64) << np.uint64(32)) | pr[:, 1].astype(np.uint64)).view(np.int64))
    d = [cu(rowptr), cu(colidx), cu(val)]
    cval, n2 = stages.center_rows(d[0], d[2])
    staged = stages.pair_centered_cosine(*d, cval, n2, packed, m, grouped=True)
    search = stages.pair_centered_cosine(*d, cval, n2, packed, m, grouped=True, force_search=True)
    shuffled = torch.randperm(packed.numel(), device=DEV)
    anyorder = stages.pair_centered_cosine(*d, cval, n2, packed[shuffled].contiguous(), m, grouped=False)
    assert np.array_equal(staged.cpu().numpy(), want)
    assert np.array_equal(search.cpu().numpy(), want)
    assert np.array_equal(anyorder.cpu().numpy(), want[shuffled.cpu().numpy()])
    if m == 300:
        assert (want > 0).sum() > 1000 and (want == 0).sum() > 1000


def test_pair_centered_cosine_long_rows():
    """Rows of ~2000 cells: the four-deep unrolled sweep of the staged kernel and its tail."""
    from query_recommendation_system_b200 import stages
    n, m = 40, 5000
    rowptr, colidx, val = _ratings_csr(n, m, 0.4, seed=8)
    ii, jj = np.triu_indices(n, 1)
    pr = np.stack([ii, jj], 1)
    want = spec.centered_cosine_milli(rowptr, colidx, val, pr)
    packed = cu(((pr[:, 0].astype(np.uint64) << np.uint64(32)) | pr[:, 1].astype(np.uint64)).view(np.int64))
    d = [cu(rowptr), cu(colidx), cu(val)]
    cval, n2 = stages.center_rows(d[0], d[2])
    got = stages.pair_centered_cosine(*d, cval, n2, packed, m, grouped=True)
    assert np.array_equal(got.cpu().numpy(), want)


def test_pair_centered_cosine_reference_user_table_c1(golden_c1):
    """The user neighbour table of the unmodified reference (recommender.py:256-286 on C1): every listed
    similarity is reproduced by the kernel for that pair."""
    from query_recommendation_system_b200 import stages
    g = golden_c1
    R = g["ratings"].astype(np.int64)
    rows, cols = np.nonzero(R)
    rowptr = np.zeros(R.shape[0] + 1, np.int64)
    np.cumsum(np.bincount(rows, minlength=R.shape[0]), out=rowptr[1:])
    pr = np.array([(i, int(g["u_idx"][i, t])) for i in range(R.shape[0]) for t in range(int(g["u_cnt"][i]))
                   if int(g["u_idx"][i, t]) != i])
    want = np.array([g["u_val"][i, t] for i in range(R.shape[0]) for t in range(int(g["u_cnt"][i]))
                     if int(g["u_idx"][i, t]) != i])
    packed = cu(((pr[:, 0].astype(np.uint64) << np.uint64(32)) | pr[:, 1].astype(np.uint64)).view(np.int64))
    d = [cu(rowptr), cu(cols.astype(np.int32)), cu(R[rows, cols].astype(np.uint8))]
    cval, n2 = stages.center_rows(d[0], d[2])
    got = stages.pair_centered_cosine(*d, cval, n2, packed, R.shape[1], grouped=True)
    assert np.array_equal(got.cpu().numpy(), np.rint(want * 1000).astype(np.int32)) and pr.shape[0] > 700


def test_pipeline_ratings_similarity_matches_oracle():
    """LshCfPipeline with similarity="ratings": LSH candidates, centred-cosine neighbour tables, predictions."""
    from query_recommendation_system_b200 import synth
    from query_recommendation_system_b200.pipeline import LshCfPipeline
    n, m, K, b, k = 1024, 2000, 64, 16, 17
    tab = spec.permutation_table(m, K, seed=7)
    rowptr, colidx, val = synth.planted_csr(n, m, 0.05, seed=21)
    pipe = LshCfPipeline(m=m, K=K, b=b, k=k, perm_tab=tab, device=DEV, distributed=False)
    pipe.similarity = "ratings"
    out = pipe.run(rowptr.to(DEV), colidx.to(DEV), val.to(DEV))
    sig = spec.signatures(rowptr.numpy(), colidx.numpy(), tab)
    cand = np.array(sorted(spec.candidates(sig, b)), dtype=np.int64).reshape(-1, 2)
    want = spec.centered_cosine_milli(rowptr.numpy(), colidx.numpy(), val.numpy(), cand)
    got_pairs = unpack(out.pairs)
    order = np.lexsort((got_pairs[:, 1], got_pairs[:, 0]))
    assert np.array_equal(got_pairs[order], cand) and np.array_equal(out.sims.cpu().numpy()[order], want)
    idx, v, cnt = spec.topk_tables(n, cand, want / 1000.0, k)
    assert np.array_equal(out.nbr_idx.cpu().numpy(), idx) and np.array_equal(out.nbr_val.cpu().numpy(), v)
    assert np.array_equal(out.nbr_cnt.cpu().numpy(), cnt)
    rows = np.arange(0, n, 37)
    ref = spec.predict_user_branch_csr(rowptr.numpy(), colidx.numpy(), val.numpy(), m, idx, v, cnt, rows=rows)
    rs, rc = out.pred_row_start.cpu().numpy(), out.pred_row_cnt.cpu().numpy()
    pc, pv = out.pred_cols.cpu().numpy(), out.pred_vals.cpu().numpy()
    assert (v[cnt > 0, 0] > 0.3).any() and (v == 0).any()
    for u in rows:
        gc = np.concatenate([pc[rs[u, x]:rs[u, x] + rc[u, x]] for x in range(rs.shape[1])])
        gp = np.concatenate([pv[rs[u, x]:rs[u, x] + rc[u, x]] for x in range(rs.shape[1])])
        rcols, rpreds = ref[u]
        assert np.array_equal(gc, rcols) and np.array_equal(gp.astype(np.int64), rpreds), u
