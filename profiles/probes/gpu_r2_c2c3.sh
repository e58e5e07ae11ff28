#!/bin/bash
# GPU suite + default bench (C2) + the C3 line on one GPU.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
t0=$SECONDS
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/t_all.log 2>&1
echo "all gpu tests rc=$? t=$((SECONDS-t0))"; tail -4 gpurun_out/t_all.log
timeout 240 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench.err
echo "bench rc=$? t=$((SECONDS-t0))"; python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/bench_c2.json'))
    print(d['ms_per_step'], d['stages']['ms'], d['e2e']['ms_per_step'], d['parity_check']['identical'])
except Exception as e:
    print('bench line unreadable', e)
PY
tail -2 gpurun_out/bench.err
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
