#!/bin/bash
# One gpurun call: MinHash / banding / drop-in parity tests, then the mode probe at C2 (and C3 when asked).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_stages.py tests/test_gpu_dropin.py -m gpu -x -q 2>&1 | tail -5
timeout 120 python profiles/probes/minhash_modes_probe.py 2>&1 | tail -2
if [ -n "$WITH_C3" ]; then
  QRS_CONFIG=C3 QRS_PROBE_N=200000 timeout 200 python profiles/probes/minhash_modes_probe.py 2>&1 | tail -2
  QRS_MINHASH_TILE_SETS=8 QRS_CONFIG=C3 QRS_PROBE_N=200000 timeout 200 python profiles/probes/minhash_modes_probe.py 2>&1 | tail -2
fi
