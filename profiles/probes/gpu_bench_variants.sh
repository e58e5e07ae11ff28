#!/bin/bash
# One gpurun call: for the in-tree library and every library under query_recommendation_system_b200/variants/
# (built with `make -C query_recommendation_system_b200/csrc variant NAME=... DEFS=...`): the MinHash + predictor
# parity tests, then the default bench line.  Outputs → gpurun_out/bench_<name>.json, gpurun_out/t_<name>.log.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
t0=$SECONDS
run_one() {   # name, library path ("" = in-tree)
  local name=$1 lib=$2
  if [ -n "$lib" ]; then export QRS_B200_LIB="$lib"; else unset QRS_B200_LIB; fi
  timeout 200 python -m pytest tests/test_gpu_stages.py -m gpu -x -q -k "minhash or predict" > "gpurun_out/t_$name.log" 2>&1
  echo "$name tests rc=$? t=$((SECONDS-t0))"; tail -1 "gpurun_out/t_$name.log"
  timeout 200 python bench.py --no-cpu-baseline > "gpurun_out/bench_$name.json" 2> "gpurun_out/bench_$name.err"
  echo "$name bench rc=$? t=$((SECONDS-t0))"
  python - "$name" <<'PY'
import json, sys
try:
    d = json.load(open('gpurun_out/bench_%s.json' % sys.argv[1]))
    print(sys.argv[1], d['ms_per_step'], d['stages']['ms'], d['e2e']['ms_per_step'])
except Exception as e:
    print(sys.argv[1], 'bench line unreadable', e)
PY
}
run_one default ""
for v in query_recommendation_system_b200/variants/*.so; do
  [ -e "$v" ] || continue
  run_one "$(basename "$v" .so)" "$PWD/$v"
done
