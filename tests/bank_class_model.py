"""NumPy model of the two layouts of a prepared (row, window) sub-segment (csrc/predict.cu: pack_rows_kernel /
pack_rows_dealt_kernel) and of what they cost the predictor: an 8-byte shared-memory access of a warp is served
half a warp at a time, one wavefront per DISTINCT address that shares a pair of banks (column mod 16), so a
half-step of 16 words costs max over bank pairs of the number of different columns in it.  Test infrastructure:
it documents the figures DESIGN.md quotes next to the ncu counts; nothing in the package imports it."""
import numpy as np

PAD = -1


def column_order(cols, steps):
    """Words in ascending column order, padded to whole steps; padding words all point at ONE spare slot
    (bank pair 0)."""
    cells = np.full(steps * 32, PAD, np.int64)
    cells[:len(cols)] = np.sort(cols)
    return cells


def bank_class_cells(cols, steps):
    """The i-th entry of class c = column mod 16 in half-step i at lane c; classes with more entries than
    half-steps park the rest ('extras') in cells left free, last step first; padding in lane l points at spare
    slot l mod 16 (a bank pair no entry of that half-step uses)."""
    H = 2 * steps
    grid = np.full((H, 16), PAD, np.int64)
    seen = np.zeros(16, np.int64)
    extras = []
    for x in np.sort(cols):
        c = int(x) % 16
        if seen[c] < H:
            grid[seen[c], c] = x
        else:
            extras.append(int(x))
        seen[c] += 1
    free = [(h, l) for s in range(steps - 1, -1, -1) for h in (2 * s + 1, 2 * s) for l in range(16) if grid[h, l] == PAD]
    for x, (h, l) in zip(extras, free):
        grid[h, l] = x
    return grid.reshape(-1), len(extras)


def wavefronts(cells, per_lane_spare):
    """Shared-memory wavefronts of one 8-byte access by the warps that hold these words, summed over half-steps."""
    total = 0
    for h in range(len(cells) // 16):
        part = cells[h * 16:(h + 1) * 16]
        real = part[part != PAD]
        load = np.bincount(real % 16, minlength=16)
        if not per_lane_spare and len(real) < 16:
            load[0] += 1                      # every padding word: the same address in bank pair 0
        total += max(1, int(load.max()))
    return total
