import os, sys, time, torch, numpy as np
import torch.distributed as dist
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
from query_recommendation_system_b200 import synth
from query_recommendation_system_b200.pipeline import LshCfPipeline
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n, m, K, b, k = 100000, 10000, 128, 32, 28
np.random.seed(7); tab = np.stack([np.random.permutation(m) for _ in range(K)], axis=1).astype(np.int32)
rowptr, colidx, val = synth.planted_csr(n, m, 0.05, seed=77, row_base=rank * n, device=dev)
pipe = LshCfPipeline(m=m, K=K, b=b, k=k, perm_tab=tab, device=dev)
for _ in range(3): out = pipe.run(rowptr, colidx, val)
idx, cnt = out.nbr_idx, out.nbr_cnt
def T(label, fn, reps=5):
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): r = fn()
    torch.cuda.synchronize()
    if rank == 0: print("%-30s %.3f ms" % (label, (time.perf_counter() - t0) * 1e3 / reps))
    return r
row_base = rank * n
live = T("live", lambda: torch.arange(k, device=dev)[None, :] < cnt[:, None])
ids = T("ids", lambda: idx.to(torch.int64))
is_remote = T("is_remote", lambda: live & ((ids < row_base) | (ids >= row_base + n)))
sel = T("ids[is_remote]", lambda: ids[is_remote])
need = T("unique", lambda: torch.unique(sel))
def ar():
    t = torch.tensor([need.numel()], dtype=torch.int64, device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); return int(t.item())
T("all_reduce+item", ar)
T("where", lambda: torch.where(live, ids - row_base, torch.full_like(ids, -1)).to(torch.int32).contiguous())
T("fetch total", lambda: pipe._fetch_remote_rows(rowptr, colidx, val, idx, cnt, n, row_base))
T("route_pairs", lambda: pipe._route_pairs(out.pairs, out.sims, n))
dist.destroy_process_group()
