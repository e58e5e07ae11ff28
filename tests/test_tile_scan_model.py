"""CPU: the tile-scan algorithm of minhash_tile_kernel (tests/tile_scan_model.py, a lane-level model of the kernel)
reproduces the golden signatures of the unmodified reference and the oracle on ragged inputs.  The CUDA kernel itself
is tested on the GPU (tests/test_gpu_stages.py::test_minhash_*)."""
import numpy as np
import pytest

from oracle import spec
from tile_scan_model import tile_scan


@pytest.mark.parametrize("force_tile", [True, False])
def test_tile_scan_model_reproduces_reference_signatures_c1(golden_c1, force_tile):
    g = golden_c1
    tab = g["perm_tab"].astype(np.int64)
    got = tile_scan(g["rowptr"], g["colidx"], tab.shape[0], tab, R=96, force_tile=force_tile)
    assert np.array_equal(got, g["sig"].astype(np.int64))


@pytest.mark.parametrize("name", ["synth_a", "synth_b"])
def test_tile_scan_model_reproduces_reference_signatures_synth(name):
    from conftest import load_golden
    g = load_golden(name)
    tab = spec.permutation_table(int(g["m"]), int(g["K"]), seed=int(g["perm_seed"])).astype(np.int64)
    n_head = 64                                          # a few tiles are enough for the model (pure Python)
    rowptr = g["rowptr"][:n_head + 1]
    got = tile_scan(rowptr, g["colidx"], int(g["m"]), tab, R=64, force_tile=True)
    assert np.array_equal(got, g["sig"][:n_head].astype(np.int64))           # includes empty sets -> -1


@pytest.mark.parametrize("n,m,K,R,force_tile,maxlen", [(70, 300, 40, 32, True, 290), (45, 500, 64, 64, False, 80),
                                                       (33, 200, 32, 96, True, 30), (64, 1000, 96, 32, False, 10),
                                                       (40, 1200, 32, 576, True, 6)])
def test_tile_scan_model_ragged_against_oracle(n, m, K, R, force_tile, maxlen):
    """partial tiles, K not a multiple of 32, sets longer than one 256-entry build step, a 32-rank head that most
    hashes of the short sets exhaust (gather-min fallback), AUTO's short sets beside scanned ones, a 576-rank head
    walked almost to its end (chunk digits above the static ones, nine high planes)"""
    rng = np.random.default_rng(n + K)
    rows = [np.sort(rng.choice(m, size=rng.integers(0, maxlen), replace=False)).astype(np.int32) for _ in range(n)]
    rowptr = np.zeros(n + 1, np.int64)
    rowptr[1:] = np.cumsum([r.size for r in rows])
    colidx = np.concatenate(rows).astype(np.int32)
    tab = spec.permutation_table(m, K, seed=3)
    ref = spec.signatures(rowptr, colidx, tab)
    assert np.array_equal(tile_scan(rowptr, colidx, m, tab.astype(np.int64), R, force_tile), ref)
