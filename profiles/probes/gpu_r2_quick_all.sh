#!/bin/bash
# One 1-GPU call: the whole GPU suite and the default bench line.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
t0=$SECONDS
timeout 500 python -m pytest tests -m gpu -x -q > gpurun_out/t_all.log 2>&1
echo "all gpu tests rc=$? t=$((SECONDS-t0))"; tail -4 gpurun_out/t_all.log
timeout 300 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench.err
echo "bench rc=$? t=$((SECONDS-t0))"; python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/bench_c2.json'))
    print(d['ms_per_step'], d['stages']['ms'], d['e2e']['ms_per_step'], d['parity_check']['identical'], d['roofline']['frac'])
except Exception as e:
    print('bench line unreadable', e)
PY
tail -2 gpurun_out/bench.err
