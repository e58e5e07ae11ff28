import os, sys, time, json, torch, numpy as np
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import bench
from query_recommendation_system_b200 import synth
from query_recommendation_system_b200.pipeline import HostRunner, LshCfPipeline
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
n, m, density, K, b, k = synth.config_shape("C2")
rowptr, colidx, val = synth.planted_csr(n, m, density, seed=bench.DATA_SEED, device=dev)
tab = bench.perm_table(m, K)
h = [x.cpu().pin_memory() for x in (rowptr, colidx, val)]
for chunks in (2, 4, 8, 16):
    pipe = LshCfPipeline(m=m, K=K, b=b, k=k, perm_tab=tab, device=dev)
    runner = HostRunner(pipe, chunks=chunks)
    for _ in range(4): runner(*h)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(6): runner(*h)
    torch.cuda.synchronize()
    print("chunks", chunks, "e2e ms", (time.perf_counter() - t0) * 1e3 / 6, "d2h MB", runner.d2h_bytes / 1e6)
