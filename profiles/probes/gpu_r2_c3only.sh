#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
t0=$SECONDS
timeout 1200 python bench.py --config C3 --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
echo "bench C3 rc=$? t=$((SECONDS-t0))"; python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/bench_c3.json'))
    print(d['ms_per_step'], d['stages']['ms'], d['e2e']['ms_per_step'], d['parity_check'].get('identical'))
except Exception as e:
    print('bench line unreadable', e)
PY
tail -2 gpurun_out/bench_c3.err
