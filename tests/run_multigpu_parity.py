#!/usr/bin/env python
"""Multi-GPU invariance check on real GPUs: run under torchrun with N ranks,

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29517 tests/run_multigpu_parity.py [--n-per-rank 20000]

Every rank runs the sharded pipeline (NCCL / peer-memory exchanges); rank 0 additionally runs the same
matrix on one GPU and checks that the union of scored pairs, the neighbour tables and the predictions
of every shard are identical (query_recommendation_system_b200/selfcheck.py).  Prints one JSON line;
exit code 1 on a mismatch.  tests/test_gpu_multigpu.py launches it when two GPUs are visible."""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-per-rank", type=int, default=20000)
    ap.add_argument("--m", type=int, default=4000)
    ap.add_argument("--density", type=float, default=0.05)
    args = ap.parse_args()
    from query_recommendation_system_b200 import selfcheck
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
    K, b, k = 128, 32, 24
    np.random.seed(7)
    tab = np.stack([np.random.permutation(args.m) for _ in range(K)], axis=1).astype(np.int32)
    res = selfcheck.multi_rank_invariance(args.n_per_rank, args.m, args.density, K, b, k, tab, dev)
    if rank == 0:
        print(json.dumps(res))
    dist.destroy_process_group()
    if rank == 0 and not res["identical"]:
        sys.exit(1)


if __name__ == "__main__":
    main()
