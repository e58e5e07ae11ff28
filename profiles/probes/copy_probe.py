"""Copy-only probe behind the multi-GPU e2e number: every rank moves, per "step", the bytes the
host-facing call moves at C2 (H2D: 16-bit column ids + ratings + row pointers = 149 MB; D2H: predictions + bitmap +
tables = 419 MB) between its own pinned buffers and its GPU — no kernels — all ranks at once.  The max over ranks is
the host-side ceiling of the e2e step at that world size.
    torchrun --nproc-per-node N profiles/probes/copy_probe.py"""
import json
import os
import time

import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
H2D, D2H = 149_439_745, 419_501_412
h_in = torch.empty(H2D, dtype=torch.uint8).pin_memory()
h_out = torch.empty(D2H, dtype=torch.uint8).pin_memory()
d_in = torch.empty(H2D, dtype=torch.uint8, device=dev)
d_out = torch.empty(D2H, dtype=torch.uint8, device=dev)
up, down = torch.cuda.Stream(), torch.cuda.Stream()


def step(overlap):
    with torch.cuda.stream(up):
        d_in.copy_(h_in, non_blocking=True)
    if not overlap:
        down.wait_stream(up)
    with torch.cuda.stream(down):
        for c in range(8):                       # the predictions come back in 8 chunks
            lo, hi = D2H * c // 8, D2H * (c + 1) // 8
            h_out[lo:hi].copy_(d_out[lo:hi], non_blocking=True)
    up.synchronize()
    down.synchronize()


res = {}
for overlap in (False, True):
    for _ in range(3):
        step(overlap)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        step(overlap)
    ms = (time.perf_counter() - t0) * 1e3 / 10
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res["overlapped_ms" if overlap else "serial_ms"] = float(t.item())
if rank == 0:
    res.update({"world": world, "h2d_bytes": H2D, "d2h_bytes": D2H,
                "aggregate_host_GBps_serial": world * (H2D + D2H) / (res["serial_ms"] * 1e-3) / 1e9,
                "d2h_GBps_per_gpu_overlapped": D2H / (res["overlapped_ms"] * 1e-3) / 1e9})
    print(json.dumps(res))
if world > 1:
    dist.destroy_process_group()
