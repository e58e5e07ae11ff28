#!/bin/bash
# A/B: prepared rows in bank-class layout (default) vs column order (QRS_PREDICT_PACK=column): predictor parity tests, then the C2 bench in both.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
t0=$SECONDS
timeout 400 python -m pytest tests/test_gpu_predict_shapes.py tests/test_gpu_stages.py -m gpu -x -q -k "predict or segments or tile or odd" > gpurun_out/t_pack.log 2>&1
echo "tests (class layout) rc=$? t=$((SECONDS-t0))"; tail -3 gpurun_out/t_pack.log
for mode in class column class; do
QRS_PREDICT_PACK=$mode timeout 240 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_pack_$mode.json 2> gpurun_out/bench_pack_$mode.err
echo "bench ($mode) rc=$? t=$((SECONDS-t0))"
python - <<PY
import json
d = json.load(open('gpurun_out/bench_pack_$mode.json'))
print('$mode', d['ms_per_step'], d['stages']['ms'], d['e2e']['ms_per_step'], (d.get('parity_check') or {}).get('identical'))
PY
done
