"""f-4: the on-disk format on the input side of the path.

`read_utility_matrix` stands in for main.py:69-75 (`dt.fread("utility_matrix.csv", columns=["user"] +
queriesIDs)`): the file written by the reference's generator (generator.py:59-77, 329-332) goes to the
device as raw bytes and is parsed there into the ratings CSR that LshCfPipeline / Recommender consume —
no dense users x queries frame on the host.  Only the header line and the user-id prefixes are decoded
on the host.

`format_matrix_csv` / `write_final_predictions` stand in for main.py:97-111 (`generator.write_csv(
"final_predictions", predictions.columns, rows)`): the text of the completed matrix is produced on the
device from the ratings CSR plus the sparse predictions of the path.
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
    ends = stages.csv_line_ends(text, max_lines=raw.size // (Q + 1) + 2)     # every line has >= Q + 1 bytes
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


def format_matrix_csv(query_ids, user_ids, rowptr, colidx, val, pred=None, zero_as_empty: bool = False) -> torch.Tensor:
    """The whole CSV (header + rows, "\\r\\n" line ends like csv.writer) as a uint8 tensor on the CSR's device.
    pred: (row_start, row_cnt, cols, preds) of a single-chunk run with int32/uint16 columns, or None.
    zero_as_empty: False = final_predictions.csv (0 printed), True = utility_matrix.csv (empty = not rated)."""
    dev = rowptr.device
    encoded = [u.encode("utf-8") for u in user_ids]
    if len(encoded) != rowptr.numel() - 1:
        raise QrsError("one user id per CSR row is required")
    ids_off = np.zeros(len(encoded) + 1, dtype=np.int64)
    np.cumsum([len(e) for e in encoded], out=ids_off[1:])
    ids = np.frombuffer(b"".join(encoded) or b"\0", dtype=np.uint8).copy()
    body = stages.csv_format(torch.from_numpy(ids).to(dev), torch.from_numpy(ids_off).to(dev), rowptr, colidx, val,
                             len(query_ids), pred, zero_as_empty)
    header = np.frombuffer((",".join(query_ids) + "\r\n").encode("utf-8"), dtype=np.uint8).copy()
    return torch.cat([torch.from_numpy(header).to(dev), body])


def write_final_predictions(path: str, query_ids, user_ids, rowptr, colidx, val, pred) -> int:
    """main.py:97-111: `final_predictions.csv` = given ratings + predictions, 0 where nothing could be predicted."""
    text = format_matrix_csv(query_ids, user_ids, rowptr, colidx, val, pred, zero_as_empty=False)
    text.cpu().numpy().tofile(path)
    return int(text.numel())
