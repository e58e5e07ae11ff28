"""Banding of an existing signature matrix (qrs_band_records): CUDA-event time and HBM fraction of the tile form
(bulk-copy staged signature tiles) — run once more with QRS_BAND_PLAIN=1 for the per-row kernel.
usage: python profiles/probes/band_probe.py [n] [K] [b]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from query_recommendation_system_b200 import stages  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
K = int(sys.argv[2]) if len(sys.argv) > 2 else 128
b = int(sys.argv[3]) if len(sys.argv) > 3 else 32
dev = torch.device("cuda", 0)
sig = torch.randint(0, 10000, (n, K), dtype=torch.int32, device=dev)
for _ in range(3):
    rec, n2 = stages.band_records(sig, b, 0)
times = []
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rec, n2 = stages.band_records(sig, b, 0)
    e1.record()
    torch.cuda.synchronize()
    times.append(e0.elapsed_time(e1))
times.sort()
ms = times[len(times) // 2]
bytes_ = n * K * 4 + n * b * 8 + n * 8
peak = 6553.9
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
print(json.dumps({"kernel": "band_records (%s)" % ("per-row" if os.environ.get("QRS_BAND_PLAIN") == "1" else "tile, bulk-copy staged"),
                  "n": n, "K": K, "b": b, "ms": ms, "algorithmic_bytes": bytes_, "GBps": bytes_ / (ms * 1e-3) / 1e9,
                  "hbm_frac": bytes_ / (ms * 1e-3) / 1e9 / peak,
                  "note": "includes the allocation of the output tensors by the wrapper (cached allocator)"}))
