"""Time qrs_peer_rows_stage on ONE device: two C2-sized shards whose prepared rows live on the same GPU (the 'remote'
rows are local memory here, so what is measured is the kernels' own cost without the link), 5 % of the table slots
pointing at the other shard."""
import ctypes
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from query_recommendation_system_b200 import stages, synth  # noqa: E402


class Two:
    def __init__(self, bufs, rank):
        self.world, self.rank = len(bufs), rank
        self.local, self.nbytes = bufs[rank].data_ptr(), bufs[rank].numel()
        self.table = (ctypes.c_uint64 * self.world)(*[b.data_ptr() for b in bufs])


n, m, k, tile = int(os.environ.get("QRS_PROBE_N", 100000)), 10000, 28, 1024
frac = float(os.environ.get("QRS_PROBE_REMOTE", 0.05))
dev = torch.device("cuda", 0)
shards = [synth.planted_csr(n, m, 0.05, seed=5 + r, device=dev) for r in range(2)]
need = max(stages.predict_workspace_bytes(n, s[1].numel(), m, tile, 1 << 20) for s in shards)
bufs = [torch.empty(need, dtype=torch.uint8, device=dev) for _ in range(2)]
plans = [stages.PredictPlan(*shards[r], m, tile, capacity_hint=1 << 20, peers=Two(bufs, r)) for r in range(2)]
g = torch.Generator(device=dev).manual_seed(1)
idx = torch.randint(0, n, (n, k), generator=g, device=dev, dtype=torch.int32)
remote = torch.rand((n, k), generator=g, device=dev) < frac
idx = torch.where(remote, idx + n, idx).contiguous()
cnt = torch.full((n,), k, dtype=torch.int32, device=dev)
probe = stages.PeerStage(plans[0], idx, cnt, 0)
needed = int(probe.needed.item())
keep = None
ts = []
for it in range(6):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    st = stages.PeerStage(plans[0], idx, cnt, needed, keep=keep)
    e1.record()
    torch.cuda.synchronize()
    keep = st.buf
    ts.append(e0.elapsed_time(e1))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
mask = (idx >= 0) & ((idx < 0) | (idx >= n))
tot = mask.sum().reshape(1).to(torch.int64)
e1.record()
torch.cuda.synchronize()
print(json.dumps({"n": n, "remote_refs": int(remote.sum().item()), "needed_steps": needed, "bytes": needed * 128,
                  "stage_ms": ts, "mask_ms": e0.elapsed_time(e1), "staged": bool(st.buf[8:16].view(torch.int64).item())}))
