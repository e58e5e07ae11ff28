#!/bin/bash
# GPU suite + bench, then the predictor captured at C3's shape (131 072 users x 100 000 queries, 98 windows per user).
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
QRS_CONFIG=C3 QRS_PROBE_N=131072 timeout 200 python profiles/probes/pipeline_shape_probe.py > gpurun_out/c3_shape.json 2> gpurun_out/c3_shape.err
echo "C3 shape probe rc=$? t=$((SECONDS-t0))"; cat gpurun_out/c3_shape.json
QRS_CONFIG=C3 QRS_PROBE_N=131072 timeout 300 ncu --set full --clock-control none --import-source on -k regex:predict_sparse -s 3 -c 1 -f \
  -o gpurun_out/predict_c3 python profiles/probes/pipeline_shape_probe.py > gpurun_out/ncu_c3.log 2>&1
echo "ncu predict C3 rc=$? t=$((SECONDS-t0))"
