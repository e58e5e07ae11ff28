"""f-4: the on-disk format on the input side of the path.

`read_utility_matrix` stands in for main.py:69-75 (`dt.fread("utility_matrix.csv", columns=["user"] +
queriesIDs)`): the file written by the reference's generator (generator.py:59-77, 329-332) goes to the
device as raw bytes and is parsed there into the ratings CSR that LshCfPipeline / Recommender consume —
no dense users x queries frame on the host.  Only the header line and the user-id prefixes are decoded
on the host.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np
import torch

from . import stages
from ._lib import QrsError


def parse_utility_matrix_bytes(raw: np.ndarray, device="cuda"):
    """raw: uint8[n] file content → (query_ids, user_ids, rowptr int64[U + 1], colidx int32[nnz], val uint8[nnz])
    with the three arrays on `device`."""
    raw = np.ascontiguousarray(raw, dtype=np.uint8)
    if not raw.flags.writeable:
        raw = raw.copy()
    nl = np.flatnonzero(raw == 10)[:1] if raw.size < (1 << 22) else _first_newline(raw)
    if nl.size == 0:
        raise QrsError("utility matrix CSV: no header line")
    header = raw[:int(nl[0])].tobytes().decode("utf-8").rstrip("\r")
    query_ids = header.split(",")
    Q = len(query_ids)
    host = torch.from_numpy(raw)
    text = (host.pin_memory() if torch.cuda.is_available() else host).to(device, non_blocking=True)
    ends = stages.csv_line_ends(text)
    n_rows = ends.numel() - 1
    rowptr, colidx, val, id_end = stages.csv_ratings(text, ends, 1, n_rows, Q)
    begins = (ends[:-1] + 1).cpu().numpy()
    stops = id_end.cpu().numpy()
    user_ids = [raw[b:e].tobytes().decode("utf-8") for b, e in zip(begins, stops)]
    return query_ids, user_ids, rowptr, colidx, val


def _first_newline(raw: np.ndarray) -> np.ndarray:
    step = 1 << 20
    for s in range(0, raw.size, step):
        hit = np.flatnonzero(raw[s:s + step] == 10)
        if hit.size:
            return hit[:1] + s
    return np.zeros(0, dtype=np.int64)


def read_utility_matrix(path: str, device="cuda") -> Tuple[List[str], List[str], torch.Tensor, torch.Tensor, torch.Tensor]:
    return parse_utility_matrix_bytes(np.fromfile(path, dtype=np.uint8), device)
