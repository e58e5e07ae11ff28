"""Run the device-resident pipeline a few times on a shard of a named config (QRS_CONFIG, QRS_PROBE_N rows) — the
process ncu attaches to when a kernel is to be captured at a shape whose full size is too slow to replay."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from query_recommendation_system_b200 import synth  # noqa: E402
from query_recommendation_system_b200.pipeline import LshCfPipeline  # noqa: E402

cfg = os.environ.get("QRS_CONFIG", "C3")
n, m, density, K, b, k = synth.config_shape(cfg)
n = int(os.environ.get("QRS_PROBE_N", n))
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
rowptr, colidx, val = synth.planted_csr(n, m, density, seed=bench.DATA_SEED, device=dev)
pipe = LshCfPipeline(m=m, K=K, b=b, k=k, perm_tab=bench.perm_table(m, K), device=dev, profile=True, distributed=False)
pipe.cols_mode = 1 if m <= 65536 else 0
for _ in range(int(os.environ.get("QRS_PROBE_REPS", 4))):
    out = pipe.run(rowptr, colidx, val)
print(json.dumps({"config": cfg, "n": n, "stages_ms": out.stage_ms, "predictions": out.n_predictions}))
