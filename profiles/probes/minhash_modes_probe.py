"""Probe: the per-set first-touch scan (mode 2) against the 32-set tile scan (mode 3 / AUTO) at a bench
config's size — equality of signatures, norm2 and bucket records, and CUDA-event times per launch."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from query_recommendation_system_b200 import _lib, stages, synth  # noqa: E402

n, m, density, K, b, k = synth.config_shape(os.environ.get("QRS_CONFIG", "C2"))
n = int(os.environ.get("QRS_PROBE_N", n))
dev = torch.device("cuda:0")
rowptr, colidx, val = synth.planted_csr(n, m, density, seed=20260102, device=dev)
np.random.seed(7)
tab = torch.from_numpy(np.stack([np.random.permutation(m) for _ in range(K)], axis=1).astype(np.int32)).to(dev)
R_probe = int(os.environ.get("QRS_PROBE_R", 0)) or stages.choose_head_ranks(m, colidx.numel() / n)
head = stages.perm_head(tab, R_probe)
REPS = int(os.environ.get("QRS_PROBE_REPS", 10))
res = {"config": [n, m, density, K, b], "R": int(head.shape[0]) * 8, "ms": {}}
outs = {}
for mode, name in ((_lib.MINHASH_FIRST_TOUCH, "first_touch"), (_lib.MINHASH_TILE, "tile"), (_lib.MINHASH_AUTO, "auto"),
                   (_lib.MINHASH_DIRECT, "direct")):
    try:
        for _ in range(2):
            o = stages.minhash(rowptr, colidx, m, tab, head, b=b, mode=mode)
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(REPS):
            o = stages.minhash(rowptr, colidx, m, tab, head, b=b, mode=mode)
        t1.record()
        torch.cuda.synchronize()
        res["ms"][name] = t0.elapsed_time(t1) / REPS
        outs[name] = o
    except _lib.QrsError as e:
        res["ms"][name] = str(e)
ref = outs.get("first_touch")
for name, o in outs.items():
    if ref is not None and name != "first_touch":
        res["equal_" + name] = bool(torch.equal(o[0], ref[0]) and torch.equal(o[1], ref[1]) and torch.equal(o[2], ref[2]))
print(json.dumps(res))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "minhash_modes_%s.jsonl" % os.environ.get("QRS_CONFIG", "C2")), "a") as f:
    f.write(json.dumps(res) + "\n")
