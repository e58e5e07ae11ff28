"""Multi-GPU invariance on hardware: candidate set, similarities, neighbour tables and
predictions of the sharded pipeline (NCCL / peer-memory exchanges) equal a single-GPU run of the same
matrix, for every power-of-two world size the box offers.  Skipped below two visible GPUs (bench.py
--gpus N runs the same check inside every scaling run and reports it as `parity_check`)."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.timeout(900)
def test_sharded_pipeline_equals_single_gpu(world):
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d visible GPUs, %d here" % (world, torch.cuda.device_count()))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "run_multigpu_parity.py"), "--n-per-rank", "12000"]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=850, cwd=ROOT)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-4000:]
    line = [ln for ln in p.stdout.splitlines() if ln.startswith("{")][-1]
    res = json.loads(line)
    assert res["identical"] and res["world"] == world and res["pairs"] > 0, res
    assert all(x > 0 for x in res["pairs_scored_per_rank"]), "the bucket exchange left a rank without work"
