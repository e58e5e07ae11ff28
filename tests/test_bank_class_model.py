"""The bank-class layout of the prepared rating rows: invariants of the cell assignment and
the shared-memory wavefront counts the design quotes (about 2.6 -> 1.67 per half-step at C2's ~50 entries per
window, 1.7 -> 1.19 at C3's ~10) — the model's column-order figure is what ncu counted on the GPU
(profiles/r2_predict_ncu.txt: 430 M conflict wavefronts of 801 M)."""
import numpy as np
import pytest

from bank_class_model import PAD, bank_class_cells, column_order, wavefronts


def windows(mean, trials, seed):
    rng = np.random.default_rng(seed)
    for _ in range(trials):
        n = int(rng.poisson(mean))
        if 0 < n <= 128:
            yield rng.choice(1024, size=n, replace=False)


def test_every_entry_once_and_one_regular_entry_per_class_and_half_step():
    for cols in windows(50, 300, 1):
        steps = (len(cols) + 31) // 32
        cells, n_extra = bank_class_cells(cols, steps)
        assert sorted(cells[cells != PAD]) == sorted(cols)
        H = 2 * steps
        counts = np.bincount(cols % 16, minlength=16)
        assert n_extra == int(np.maximum(counts - H, 0).sum())
        grid = cells.reshape(H, 16)
        for c in range(16):                    # lane c of the first min(count, H) half-steps holds class c
            own = grid[:min(counts[c], H), c]
            assert np.all(own % 16 == c) and np.all(own != PAD)


def test_single_class_window_parks_everything_beyond_the_half_steps():
    cols = 16 * np.arange(40)                  # one class: 40 entries, 2 steps = 4 half-steps
    cells, n_extra = bank_class_cells(cols, 2)
    assert n_extra == 36 and sorted(cells[cells != PAD]) == sorted(cols)


@pytest.mark.parametrize("mean,plain,dealt", [(50, 2.55, 1.67), (10, 1.70, 1.19)])
def test_wavefronts_per_half_step(mean, plain, dealt):
    a = b = halves = 0
    for cols in windows(mean, 2500, 7):
        steps = (len(cols) + 31) // 32
        a += wavefronts(column_order(cols, steps), per_lane_spare=False)
        b += wavefronts(bank_class_cells(cols, steps)[0], per_lane_spare=True)
        halves += 2 * steps
    assert abs(a / halves - plain) < 0.08, a / halves
    assert abs(b / halves - dealt) < 0.08, b / halves
