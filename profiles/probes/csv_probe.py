"""Probe: one format + parse round trip of the C2 utility matrix (for `ncu --metrics gpu__time_duration.sum`)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from query_recommendation_system_b200 import stages, synth  # noqa: E402

n, m, density, K, b, k = synth.config_shape(os.environ.get("QRS_CONFIG", "C2"))
dev = torch.device("cuda:0")
rowptr, colidx, val = synth.planted_csr(n, m, density, seed=20260102, device=dev)
enc = [("U%d" % (i + 1)).encode() for i in range(n)]
ids_off = torch.from_numpy(np.concatenate([[0], np.cumsum([len(e) for e in enc])]).astype(np.int64)).to(dev)
ids = torch.from_numpy(np.frombuffer(b"".join(enc), dtype=np.uint8).copy()).to(dev)
header = torch.from_numpy(np.frombuffer((",".join("Q%d" % (j + 1) for j in range(m)) + "\r\n").encode(), dtype=np.uint8).copy()).to(dev)
for rep in range(2):
    body = stages.csv_format(ids, ids_off, rowptr, colidx, val, m, None, True)
    text = torch.cat([header, body])
    ends = stages.csv_line_ends(text, max_lines=text.numel() // (m + 1) + 2)
    rp, ci, v, _ = stages.csv_ratings(text, ends, 1, ends.numel() - 1, m)
    torch.cuda.synchronize()
print("identical", bool(torch.equal(rp, rowptr) and torch.equal(ci, colidx) and torch.equal(v, val)), text.numel())
