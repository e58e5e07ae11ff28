#!/bin/bash
# A/B: prepared rows in bank-class layout (default at C2) vs column order (QRS_PREDICT_PACK=column): predictor parity tests, then the C2 bench in both; optional variant libraries of the row preparation.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
t0=$SECONDS
timeout 400 python -m pytest tests/test_gpu_predict_shapes.py tests/test_gpu_stages.py -m gpu -x -q -k "predict or segments or tile or odd or layout" > gpurun_out/t_pack.log 2>&1
echo "tests rc=$? t=$((SECONDS-t0))"; tail -3 gpurun_out/t_pack.log
QRS_PREDICT_PACK=class timeout 400 python -m pytest tests/test_gpu_predict_shapes.py tests/test_gpu_stages.py -m gpu -x -q -k "predict or segments or tile or odd" > gpurun_out/t_pack_class.log 2>&1
echo "tests (class layout forced) rc=$? t=$((SECONDS-t0))"; tail -3 gpurun_out/t_pack_class.log
show() { python - <<PY
import json
d = json.load(open('gpurun_out/bench_pack_$1.json'))
print('$1', d['ms_per_step'], d['stages']['ms'], d['e2e']['ms_per_step'], (d.get('parity_check') or {}).get('identical'))
PY
}
for mode in class column class; do
QRS_PREDICT_PACK=$mode timeout 240 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_pack_$mode.json 2> gpurun_out/bench_pack_$mode.err
echo "bench ($mode) rc=$? t=$((SECONDS-t0))"; show $mode
done
for v in query_recommendation_system_b200/variants/libqrs_*.so; do
[ -f "$v" ] || continue
name=$(basename $v .so)
QRS_B200_LIB=$PWD/$v timeout 240 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_pack_$name.json 2> gpurun_out/bench_pack_$name.err
echo "bench ($name) rc=$? t=$((SECONDS-t0))"; show $name
done
