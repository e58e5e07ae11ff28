import os, sys, time, json, torch, numpy as np
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import bench
from query_recommendation_system_b200 import synth
from query_recommendation_system_b200.pipeline import HostRunner, LshCfPipeline
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
n, m, density, K, b, k = synth.config_shape("C2")
rowptr, colidx, val = synth.planted_csr(n, m, density, seed=bench.DATA_SEED, device=dev)
tab = bench.perm_table(m, K)
h_rowptr, h_val = rowptr.cpu().pin_memory(), val.cpu().pin_memory()
h_colidx = torch.from_numpy(colidx.cpu().numpy().astype(np.uint16)).pin_memory()      # the compact wire format bench.py uses
h = [h_rowptr, h_colidx, h_val]
for schedule, chunks, growth in (("equal", 8, 1.0), ("ramp", 6, 1.7), ("ramp", 8, 1.7), ("ramp", 8, 1.5), ("ramp", 10, 1.5), ("ramp", 12, 1.4)):
    pipe = LshCfPipeline(m=m, K=K, b=b, k=k, perm_tab=tab, device=dev)
    pipe.chunk_growth = growth
    runner = HostRunner(pipe, chunks=chunks, schedule=schedule)
    for _ in range(4): runner(*h)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(8): runner(*h)
    torch.cuda.synchronize()
    print(json.dumps({"schedule": schedule, "chunks": chunks, "growth": growth, "e2e_ms": (time.perf_counter() - t0) * 1e3 / 8,
                      "d2h_MB": runner.d2h_bytes / 1e6, "h2d_MB": runner.h2d_bytes / 1e6}))
