"""Synthetic utility matrices in the schema of the reference's resources/generator.py
(integer ratings 1..100, 0/empty = missing; generator.py:19, 264-334), scaled to the
BASELINE.json shapes with the planted-taste-group recipe of SURVEY.md §8(d).

Everything is derived from counter-based integer hashes (splitmix64 on int64 tensors),
so the same (shape, seed) gives bit-identical data on CPU and on CUDA and any row
range can be regenerated independently (rank r of a multi-GPU run builds only its shard).

Recipe: groups of `group` consecutive sets share a prototype of round(density*m)
elements; each member keeps a prototype element w.p. 0.8 and adds Binomial(base, 0.2)
uniform-random elements (within-group Jaccard ~0.47); rating = clip(mu(group, element) +
noise(sd 10), 1, 100) with mu ~ U[20, 90].
"""
from __future__ import annotations

import torch

_M1 = -7046029254386353131            # 0x9E3779B97F4A7C15 as int64
_M2 = -4658895280553007687            # 0xBF58476D1CE4E5B9
_M3 = -7723592293110705685            # 0x94D049BB133111EB


def _lsr(x: torch.Tensor, s: int) -> torch.Tensor:
    return (x >> s) & ((1 << (64 - s)) - 1)


def _mix(x: torch.Tensor) -> torch.Tensor:
    x = x + _M1
    x = (x ^ _lsr(x, 30)) * _M2
    x = (x ^ _lsr(x, 27)) * _M3
    return x ^ _lsr(x, 31)


def _h(seed: int, stream: int, a: torch.Tensor, b) -> torch.Tensor:
    """63-bit non-negative hash of (seed, stream, a, b)."""
    x = _mix(a * 0x632BE5AB + (seed * 1000003 + stream))
    x = _mix(x ^ (b * 0x2545F491 + 0x1B873593))
    return _lsr(x, 1)


def planted_csr(n: int, m: int, density: float, seed: int, group: int = 32, row_base: int = 0,
                device="cpu", group_chunk: int = 512):
    """CSR (rowptr int64[n+1], colidx int32[nnz] ascending per row, val uint8[nnz]) of rows
    [row_base, row_base + n) of the (conceptually unbounded) planted matrix."""
    dev = torch.device(device)
    base = max(1, round(density * m))
    g_lo = row_base // group                      # a shard may start (and end) inside a taste group
    n_groups = (row_base % group + n + group - 1) // group
    P = 1 << 20
    keys_all = []
    for c0 in range(0, n_groups, group_chunk):
        gc = min(group_chunk, n_groups - c0)
        gid = torch.arange(g_lo + c0, g_lo + c0 + gc, device=dev, dtype=torch.int64)
        # prototype = the `base` elements with the smallest hash score in this group
        elem = torch.arange(m, device=dev, dtype=torch.int64)
        score = _h(seed, 1, gid[:, None], elem[None, :])
        proto = torch.topk(score, base, dim=1, largest=False).indices          # [gc, base]
        rows = (gid[:, None] * group + torch.arange(group, device=dev, dtype=torch.int64)[None, :]).reshape(-1)
        valid_row = (rows >= row_base) & (rows < row_base + n)
        slot = torch.arange(base, device=dev, dtype=torch.int64)
        keep = (_h(seed, 2, rows[:, None], slot[None, :]) % P) < int(0.8 * P)   # [gc*group, base]
        proto_rows = proto.repeat_interleave(group, dim=0)                       # [gc*group, base]
        extra_col = _h(seed, 3, rows[:, None], slot[None, :]) % m
        extra_on = (_h(seed, 4, rows[:, None], slot[None, :]) % P) < int(0.2 * P)
        keep &= valid_row[:, None]
        extra_on &= valid_row[:, None]
        k1 = (rows[:, None] * m + proto_rows)[keep]
        k2 = (rows[:, None] * m + extra_col)[extra_on]
        keys_all.append(torch.unique(torch.cat([k1, k2])))                       # sorted (row, col), distinct
    keys = torch.cat(keys_all) if keys_all else torch.zeros(0, dtype=torch.int64, device=dev)
    rows = torch.div(keys, m, rounding_mode="floor")
    cols = keys - rows * m
    counts = torch.bincount(rows - row_base, minlength=n)
    rowptr = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    rowptr[1:] = torch.cumsum(counts, 0)
    # rating: mu(group, element) in [20, 90] + approx-normal noise (sd 10) from four 16-bit uniforms
    grp = torch.div(rows, group, rounding_mode="floor")
    mu = 20 + _h(seed, 5, grp, cols) % 71
    u = _h(seed, 6, rows, cols)
    s4 = (u & 0xFFFF) + ((u >> 16) & 0xFFFF) + ((u >> 32) & 0xFFFF) + ((u >> 47) & 0xFFFF) - 2 * 65536
    noise = torch.div(s4 * 1732, 100 * 65536, rounding_mode="trunc")
    val = torch.clamp(mu + noise, 1, 100).to(torch.uint8)
    return rowptr, cols.to(torch.int32), val


def config_shape(name: str):
    """BASELINE.json `configs` → (n sets, m elements, density, K, b, k)."""
    shapes = {
        # name: (n, m, density, K, b, k)   k = round(log_1.5 n) (recommender.py:151) unless the config fixes it
        "C2": (100_000, 10_000, 0.05, 128, 32, 28),
        "C3": (1_000_000, 100_000, 0.01, 256, 64, 34),
        "C4": (10_000_000, 1_000_000, 0.001, 128, 32, 50),
        "tiny": (2_048, 1_000, 0.05, 128, 32, 19),
    }
    return shapes[name]
