"""Probe for the "next" rows f-3 / f-4 at C2 size (for `ncu --metrics gpu__time_duration.sum -k regex:csv|center`):
one format + parse round trip of the utility matrix, and the centred-cosine scoring of the LSH candidate pairs."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from query_recommendation_system_b200 import stages, synth  # noqa: E402
from query_recommendation_system_b200.pipeline import LshCfPipeline  # noqa: E402

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

np.random.seed(7)
tab = np.stack([np.random.permutation(m) for _ in range(K)], axis=1).astype(np.int32)
pipe = LshCfPipeline(m=m, K=K, b=b, k=k, perm_tab=tab, device=dev, distributed=False)
pipe.similarity = "ratings"
for rep in range(2):
    out = pipe.run(rowptr, colidx, val, predict=False)
    torch.cuda.synchronize()
print("pairs", out.pairs.numel(), "max milli", int(out.sims.max()))
