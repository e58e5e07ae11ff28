"""Drop-in for the reference's `lsh.LSH` (lsh.py:7-54) running on the B200 kernels.

Same surface — `LSH(b)`, `make_subvecs(sig)`, `compute_buckets(sig)` once per set
in row order, `get_candidates(signatures) -> set[(i, j)]` with i < j — plus batch
entry points (`compute_buckets_batch`, `get_candidates_packed`, `candidates_scored`)
that keep everything on the device.  State is per instance (the reference's
class-level `buckets` list, lsh.py:9, is a sharing bug that is not reproduced).
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from . import stages


def _device(device=None) -> torch.device:
    if device is not None:
        return torch.device(device)
    if not torch.cuda.is_available():
        from ._lib import QrsError
        raise QrsError("query_recommendation_system_b200 needs a CUDA device; there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def group_and_pair(records: torch.Tensor, n_per_seg: int, segs: int, sig: torch.Tensor,
                   norm2: Optional[torch.Tensor], b: int, want_sims: bool, sig_id_base: int = 0,
                   bit_lo: int = 40, bit_hi: int = 64):
    """records (flat int64, `segs` segments) → unique candidate pairs (+ rounded cosines).
    Grouping = radix sort on the hash bits; pairs = slots of equal-(hash24, band) runs verified on
    the exact lanes and de-duplicated by first matching band (lsh.py:40-54)."""
    flat = records.reshape(-1)
    stages.sort_u64(flat, n_per_seg, segs, bit_lo, bit_hi)
    slot_off = stages.pair_slots(flat)
    n_slots = int(slot_off[-1].item())
    pairs, sims, counter = stages.pairs_score(flat, slot_off, sig, norm2, b, n_slots, sig_id_base, want_sims)
    n_pairs = int(counter.item())
    assert n_pairs <= n_slots
    return pairs[:n_pairs], (sims[:n_pairs] if sims is not None else None), n_slots


class LSH:
    def __init__(self, b: int, device=None):
        self.b = int(b)
        self.counter = 0
        self._rows = []                # host-buffered signatures (compute_buckets, one per set)
        self._sig = None               # device int32[n, K]
        self._device = device

    # ------------------------------------------------------------------ reference surface
    def make_subvecs(self, signature):
        """lsh.py:17-28: the signature cut into b rows of r values, wrapped to int16.
        (Pure reshaping of one row for callers that inspect it; the bucket ids the kernels use
        are produced on the device by `band_keys`.)"""
        signature = np.asarray(signature)
        l = len(signature)
        assert l % self.b == 0
        return signature.reshape(self.b, l // self.b).astype("int16")

    def compute_buckets(self, signature):
        """lsh.py:31-38: register one set's signature; its id is the call order."""
        signature = np.asarray(signature)
        assert len(signature) % self.b == 0
        self._rows.append(signature.astype(np.int64, copy=False))
        self._sig = None
        self.counter += 1

    def get_candidates(self, signatures=None) -> set:
        """lsh.py:40-54 → `set[(i, j)]`, i < j.  Materialising Python tuples is only sensible at
        reference scale; use `get_candidates_packed` otherwise."""
        packed = self.get_candidates_packed().cpu().numpy().view(np.uint64)
        return set(zip((packed >> np.uint64(32)).astype(np.int64).tolist(),
                       (packed & np.uint64(0xFFFFFFFF)).astype(np.int64).tolist()))

    # ------------------------------------------------------------------ batch / device surface
    def compute_buckets_batch(self, signatures):
        """All sets at once: `signatures` is [n, K] (numpy or CUDA tensor), ids = row numbers
        continuing from the current counter."""
        if isinstance(signatures, torch.Tensor):
            if self._rows or self._sig is not None:
                prev = self.signatures_device()
                self._sig = torch.cat([prev, signatures.to(prev.device, torch.int32)])
                self._rows = []
            else:
                self._sig = signatures.to(_device(self._device), torch.int32).contiguous()
            self.counter = self._sig.shape[0]
            assert self._sig.shape[1] % self.b == 0
        else:
            for s in np.asarray(signatures):
                self.compute_buckets(s)

    def signatures_device(self) -> torch.Tensor:
        if self._sig is None:
            dev = _device(self._device)
            if not self._rows:
                return torch.empty((0, self.b), dtype=torch.int32, device=dev)
            host = np.stack(self._rows)
            # values are permutation ranks or -1 (recommender.py:114-129): they fit int32
            assert host.min() >= -(2 ** 31) and host.max() < 2 ** 31
            self._sig = torch.from_numpy(host.astype(np.int32)).to(dev)
        return self._sig

    def band_keys(self):
        """(keys int64[n, b], keys_hi | None, skipped uint8[n, b]) — the exact bucket ids."""
        return stages.band_keys(self.signatures_device(), self.b)

    def candidates_scored(self, want_sims: bool = True):
        """(pairs int64[P] packed i<<32|j, sim_milli int32[P] | None) on the device."""
        sig = self.signatures_device()
        n = sig.shape[0]
        if n == 0:
            dev = sig.device
            return torch.empty(0, dtype=torch.int64, device=dev), torch.empty(0, dtype=torch.int32, device=dev)
        records, norm2 = stages.band_records(sig, self.b, 0, want_norm2=want_sims)
        pairs, sims, _ = group_and_pair(records, n, self.b, sig, norm2, self.b, want_sims)
        return pairs, sims

    def get_candidates_packed(self) -> torch.Tensor:
        """Sorted unique int64 (i << 32 | j), i < j, on the device."""
        pairs, _ = self.candidates_scored(want_sims=False)
        return torch.sort(pairs).values

    @property
    def buckets(self):
        """The reference's `buckets` attribute (list of b dicts, str key → [ids], lsh.py:9, 32-37),
        rebuilt on demand from the device bucket ids — introspection only."""
        keys, hi, _ = self.band_keys()
        keys = keys.cpu().numpy().view(np.uint64)
        hi = hi.cpu().numpy().view(np.uint64) if hi is not None else None
        n, b = keys.shape
        K = self.signatures_device().shape[1]
        r = K // b
        out = []
        for band in range(b):
            d = {}
            for s in range(n):
                lanes = []
                for c in range(r):
                    pos = r - 1 - c          # lane c counted from the most significant end
                    word = keys[s, band] if pos < 4 else hi[s, band]
                    v = int((int(word) >> (16 * (pos % 4))) & 0xFFFF)
                    lanes.append(str(v - 65536 if v >= 32768 else v))
                d.setdefault(",".join(lanes), []).append(s)
            out.append(d)
        return out
