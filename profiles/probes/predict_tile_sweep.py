#!/usr/bin/env python
"""Probe: predictor time at the bench shape for several window widths (one JSON line per width).
   python profiles/probes/predict_tile_sweep.py [--config C2] [--tiles 512,1024,...]"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from query_recommendation_system_b200 import _lib, stages, synth      # noqa: E402
from query_recommendation_system_b200.pipeline import LshCfPipeline  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C2")
    ap.add_argument("--tiles", default="512,768,1024,1280,1664,2016,2528,3360,5024")
    ap.add_argument("--n", type=int, default=0)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    n, m, density, K, b, k = synth.config_shape(args.config)
    n = args.n or n
    rowptr, colidx, val = synth.planted_csr(n, m, density, seed=20260102, device=dev)
    np.random.seed(7)
    tab = np.stack([np.random.permutation(m) for _ in range(K)], axis=1).astype(np.int32)
    pipe = LshCfPipeline(m=m, K=K, b=b, k=k, perm_tab=tab, device=dev)
    out = pipe.run(rowptr, colidx, val, predict=False)
    idx, v, cnt = out.nbr_idx, out.nbr_val, out.nbr_cnt
    milli = torch.round(v * 1000).to(torch.int32)
    regions = _lib.load().qrs_device_sm_count(0)
    for tile in [int(x) for x in args.tiles.split(",")]:
        warps = _lib.load().qrs_predict_warps_per_sm(tile, 1)
        if warps < 1:
            continue
        res = {"tile": tile, "nt": -(-m // tile), "warps_per_sm": warps}
        try:
            plan = stages.PredictPlan(rowptr, colidx, val, m, tile)
            _, _, _, _, demand, _ = stages.predict_sparse(rowptr, colidx, val, m, (idx, v, cnt), 0, n, tile, 0, u_milli=milli,
                                                          n_regions=regions, cols_mode=2, plan=plan)
            rcap = int(demand.max().item() * 1.03) + 256
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            tp, tq = [], []
            for it in range(4):
                ev[0].record()
                plan = stages.PredictPlan(rowptr, colidx, val, m, tile, capacity_hint=rcap * regions)
                ev[1].record()
                stages.predict_sparse(rowptr, colidx, val, m, (idx, v, cnt), 0, n, tile, rcap * regions, u_milli=milli,
                                      n_regions=regions, cols_mode=2, plan=plan)
                ev[2].record()
                torch.cuda.synchronize()
                if it:
                    tp.append(ev[0].elapsed_time(ev[1]))
                    tq.append(ev[1].elapsed_time(ev[2]))
            res.update({"prepare_ms": float(np.median(tp)), "predict_ms": float(np.median(tq))})
        except Exception as e:      # noqa: BLE001
            res["error"] = str(e)[:200]
        print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
