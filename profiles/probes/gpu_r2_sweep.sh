#!/bin/bash
# GPU suite, default bench, band/row sweeps (BASELINE configs[4]) at C3 and C2 with the current kernels.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
t0=$SECONDS
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/t_all.log 2>&1
echo "all gpu tests rc=$? t=$((SECONDS-t0))"; tail -3 gpurun_out/t_all.log
timeout 240 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench.err
echo "bench rc=$? t=$((SECONDS-t0))"; python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/bench_c2.json'))
    print(d['ms_per_step'], d['stages']['ms'], d['e2e']['ms_per_step'], d['parity_check']['identical'])
except Exception as e:
    print('bench line unreadable', e)
PY
timeout 600 python bench.py --sweep --config C3 > gpurun_out/sweep_c3.json 2> gpurun_out/sweep_c3.err
echo "sweep C3 rc=$? t=$((SECONDS-t0))"; cat gpurun_out/sweep_c3.json | cut -c1-900
timeout 300 python bench.py --sweep --config C2 > gpurun_out/sweep_c2.json 2> gpurun_out/sweep_c2.err
echo "sweep C2 rc=$? t=$((SECONDS-t0))"; cat gpurun_out/sweep_c2.json | cut -c1-900
