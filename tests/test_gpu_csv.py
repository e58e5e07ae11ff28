"""f-4 on the GPU: utility_matrix.csv (generator.py:59-77, 329-332; read by main.py:69-75) parsed on the device
into the ratings CSR, against the oracle's csv-module reading (oracle/spec.py:parse_utility_csv) and the
matrix the unmodified reference loaded from the same golden file."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import spec

pytestmark = pytest.mark.gpu
DEV = "cuda"


def csr_to_dense(rowptr, colidx, val, Q):
    rowptr, colidx, val = rowptr.cpu().numpy(), colidx.cpu().numpy(), val.cpu().numpy()
    dense = np.zeros((rowptr.size - 1, Q), dtype=np.int64)
    rows = np.repeat(np.arange(rowptr.size - 1), np.diff(rowptr))
    dense[rows, colidx] = val
    assert all((np.diff(colidx[rowptr[r]:rowptr[r + 1]]) > 0).all() for r in range(rowptr.size - 1))   # ascending
    assert (val != 0).all()
    return dense


def make_csv(dense, user_ids, query_ids, eol="\r\n", trailing=True, zero_as_text=False):
    lines = [",".join(query_ids)]
    for u, row in zip(user_ids, dense):
        lines.append(u + "," + ",".join(("0" if zero_as_text and (i % 7 == 0) else "") if v == 0 else str(int(v))
                                        for i, v in enumerate(row)))
    return (eol.join(lines) + (eol if trailing else "")).encode()


def test_golden_c1_file_matches_reference_matrix(golden_c1):
    from query_recommendation_system_b200 import io_csv
    g = golden_c1
    q, u, rowptr, colidx, val = io_csv.read_utility_matrix(os.path.join(GOLDEN, "c1", "utility_matrix.csv"), DEV)
    assert q == g["queries_ids"].tolist() and u == g["users_ids"].tolist()
    assert np.array_equal(csr_to_dense(rowptr, colidx, val, len(q)), g["ratings"].astype(np.int64))


@pytest.mark.parametrize("U,Q,fill,eol,trailing", [
    (300, 1000, 0.05, "\r\n", True), (64, 5000, 0.3, "\n", True), (33, 31, 0.9, "\r\n", False),
    (5, 1, 0.5, "\n", False), (200, 257, 0.0, "\r\n", True)])
def test_synthetic_files_match_csv_module(U, Q, fill, eol, trailing):
    from query_recommendation_system_b200 import io_csv
    rng = np.random.default_rng(U * 1000 + Q)
    dense = np.where(rng.random((U, Q)) < fill, rng.integers(1, 101, (U, Q)), 0)
    dense[rng.random((U, Q)) < fill * 0.05] = 100                       # three-digit fields
    if U > 3:
        dense[3, :] = 7                                                  # a full row
        dense[2, :] = 0                                                  # an empty row (only commas)
    users = ["U%d" % (i * 37 + 1) for i in range(U)]
    queries = ["Q%d" % (j + 1) for j in range(Q)]
    data = make_csv(dense, users, queries, eol, trailing, zero_as_text=True)
    oq, ou, od = spec.parse_utility_csv(data)
    assert oq == queries and ou == users and np.array_equal(od, dense)   # the oracle reads its own writer
    q, u, rowptr, colidx, val = io_csv.parse_utility_matrix_bytes(np.frombuffer(data, dtype=np.uint8), DEV)
    assert q == queries and u == users
    assert np.array_equal(csr_to_dense(rowptr, colidx, val, Q), dense)


def test_line_ends_across_chunk_boundaries():
    from query_recommendation_system_b200 import stages
    rng = np.random.default_rng(3)
    for n in (1, 15, 16, 17, 511, 512, 513, 4095, 4096, 4097, 100_003):
        raw = rng.integers(32, 127, n).astype(np.uint8)
        raw[raw == 10] = 32
        marks = [p for p in (0, 15, 16, 511, 512, 4095, 4096, 4097, 8191, 8192, n - 1) if p < n]
        marks += rng.integers(0, n, max(1, n // 300)).tolist()
        raw[np.array(marks)] = 10
        want = np.flatnonzero(raw == 10)
        for tail_open in (False, True):
            if tail_open:
                if n == 1:
                    continue
                raw[-1] = 65
                want = np.append(np.flatnonzero(raw == 10), n)
            t = torch.from_numpy(raw.copy()).to(DEV)
            for bound in (None, want.size + 3, 1):              # sizing pass / one pass with a bound / bound too small
                got = stages.csv_line_ends(t, max_lines=bound)
                assert np.array_equal(got.cpu().numpy(), want), (n, tail_open, bound)


@pytest.mark.parametrize("bad_row,text", [
    (1, "U2,1,2\r\n"),                 # a field short
    (1, "U2,1,2,3,4\r\n"),             # a field too many
    (1, "U2,1,x,3\r\n"),               # not a number
    (1, "U2,1,1000,3\r\n"),            # four digits
    (1, "U2,1,256,3\r\n"),             # beyond a byte
    (1, "U2,1,\"2\",3\r\n"),           # quoted
    (1, "U2,1, 2,3\r\n"),              # blank inside a field
    (1, "\r\n"),                       # empty line
])
def test_malformed_rows_are_reported(bad_row, text):
    from query_recommendation_system_b200 import io_csv
    from query_recommendation_system_b200._lib import QrsError
    data = ("Q1,Q2,Q3\r\nU1,5,,7\r\n" + text + "U3,,,\r\n").encode()
    with pytest.raises(QrsError, match="first at data row %d" % bad_row):
        io_csv.parse_utility_matrix_bytes(np.frombuffer(data, dtype=np.uint8), DEV)


def test_parsed_file_feeds_the_recommender(golden_c1):
    """main.py's sequence with the GPU reader in place of dt.fread for the utility matrix."""
    import pandas as pd
    from query_recommendation_system_b200 import io_csv
    from query_recommendation_system_b200.recommender import Recommender
    c1 = os.path.join(GOLDEN, "c1")
    q, u, rowptr, colidx, val = io_csv.read_utility_matrix(os.path.join(c1, "utility_matrix.csv"), DEV)
    rec = Recommender(verbose=False)
    dataset = pd.read_csv(os.path.join(c1, "dataset.csv"), dtype=str, keep_default_na=False)
    rec.datasetFeatures = list(dataset.columns)[1:]
    queries, queries_ids = rec.parse_queries(os.path.join(c1, "queries.csv"))
    assert queries_ids == q
    rec.init(np.array(u)[:, None], queries, queries_ids, dataset, csr_to_dense(rowptr, colidx, val, len(q)))
    assert np.array_equal(rec.ratings, golden_c1["ratings"].astype(np.int32))


# ------------------------------------------------------------------------------ export
def _csr(dense):
    rows, cols = np.nonzero(dense)
    rowptr = np.zeros(dense.shape[0] + 1, np.int64)
    np.cumsum(np.bincount(rows, minlength=dense.shape[0]), out=rowptr[1:])
    to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    return to(rowptr), to(cols.astype(np.int32)), to(dense[rows, cols].astype(np.uint8))


def test_export_reproduces_the_reference_written_file(golden_c1):
    """utility_matrix.csv of the golden set was written by the reference's generator: same bytes."""
    from query_recommendation_system_b200 import io_csv
    g = golden_c1
    text = io_csv.format_matrix_csv(g["queries_ids"].tolist(), g["users_ids"].tolist(), *_csr(g["ratings"].astype(np.int64)),
                                    zero_as_empty=True)
    raw = open(os.path.join(GOLDEN, "c1", "utility_matrix.csv"), "rb").read()
    assert text.cpu().numpy().tobytes() == raw
    # and main.py:97-111's export of the reference's own final matrix
    final = g["predictions"].astype(np.int64)
    text = io_csv.format_matrix_csv(g["queries_ids"].tolist(), g["users_ids"].tolist(), *_csr(final), zero_as_empty=False)
    assert text.cpu().numpy().tobytes() == spec.format_matrix_csv(g["queries_ids"].tolist(), g["users_ids"].tolist(),
                                                                 final, False)


@pytest.mark.parametrize("U,Q,fill", [(7, 1, 0.5), (40, 31, 0.5), (50, 4096, 0.1), (50, 4097, 0.1), (20, 9000, 0.6),
                                      (3, 70_000, 0.01)])
@pytest.mark.parametrize("zero_as_empty", [False, True])
def test_export_matches_csv_writer_and_round_trips(U, Q, fill, zero_as_empty):
    from query_recommendation_system_b200 import io_csv
    rng = np.random.default_rng(Q + U)
    dense = np.where(rng.random((U, Q)) < fill, rng.integers(1, 256, (U, Q)), 0)
    dense[rng.random((U, Q)) < fill / 4] = 100
    dense[0, :] = 0
    dense[U - 1, :] = 255
    users = ["user-%d" % (i * i) for i in range(U)]
    users[1] = ""                                                       # an empty id is still a field
    queries = ["Q%d" % (j + 1) for j in range(Q)]
    rowptr, colidx, val = _csr(dense)
    text = io_csv.format_matrix_csv(queries, users, rowptr, colidx, val, zero_as_empty=zero_as_empty)
    data = text.cpu().numpy().tobytes()
    assert data == spec.format_matrix_csv(queries, users, dense, zero_as_empty)
    q, u, rp, ci, v = io_csv.parse_utility_matrix_bytes(np.frombuffer(data, dtype=np.uint8), DEV)
    assert q == queries and u == users
    assert torch.equal(rp, rowptr) and torch.equal(ci, colidx) and torch.equal(v, val)


@pytest.mark.parametrize("cols_mode", [0, 1])
def test_export_final_predictions_from_the_pipeline(cols_mode):
    """Ratings CSR + the path's sparse predictions → final_predictions.csv, against the csv.writer oracle on the
    densely assembled matrix (main.py:97-111)."""
    from query_recommendation_system_b200 import io_csv, synth
    from query_recommendation_system_b200.pipeline import LshCfPipeline
    n, m, K, b, k = 512, 1500, 64, 16, 15
    tab = spec.permutation_table(m, K, seed=7)
    rowptr, colidx, val = [x.to(DEV) for x in synth.planted_csr(n, m, 0.05, seed=33)]
    pipe = LshCfPipeline(m=m, K=K, b=b, k=k, perm_tab=tab, device=DEV, distributed=False)
    pipe.cols_mode = cols_mode
    out = pipe.run(rowptr, colidx, val)
    users = ["U%d" % (i + 1) for i in range(n)]
    queries = ["Q%d" % (j + 1) for j in range(m)]
    pred = (out.pred_row_start, out.pred_row_cnt, out.pred_cols, out.pred_vals)
    text = io_csv.format_matrix_csv(queries, users, rowptr, colidx, val, pred)
    dense = csr_to_dense(rowptr, colidx, val, m)
    rs, rc = out.pred_row_start.cpu().numpy(), out.pred_row_cnt.cpu().numpy()
    pc, pv = out.pred_cols.cpu().numpy(), out.pred_vals.cpu().numpy()
    if cols_mode == 1:
        pc = pc.view(np.uint16)
    filled = 0
    for u in range(n):
        for t in range(rs.shape[1]):
            s, c = rs[u, t], rc[u, t]
            assert (dense[u, pc[s:s + c]] == 0).all()                    # only unrated cells are predicted
            dense[u, pc[s:s + c]] = pv[s:s + c]
            filled += c
    assert filled > 10_000
    assert text.cpu().numpy().tobytes() == spec.format_matrix_csv(queries, users, dense, False)


def test_export_rejects_too_many_columns():
    from query_recommendation_system_b200 import io_csv
    from query_recommendation_system_b200._lib import QrsError
    dense = np.zeros((1, 4), dtype=np.int64)
    rowptr, colidx, val = _csr(dense)
    with pytest.raises(QrsError):
        io_csv.format_matrix_csv(["Q"] * 200_000, ["U1"], rowptr, torch.zeros(1, dtype=torch.int32, device=DEV),
                                 torch.zeros(1, dtype=torch.uint8, device=DEV))
