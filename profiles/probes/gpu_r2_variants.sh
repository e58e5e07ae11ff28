#!/bin/bash
# A/B of predictor build variants (query_recommendation_system_b200/variants/*.so) against the in-tree library:
# the default bench line without the CPU leg and extras, parity check on.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
t0=$SECONDS
run_one() {
  local name=$1 lib=$2
  if [ -n "$lib" ]; then export QRS_B200_LIB="$lib"; else unset QRS_B200_LIB; fi
  timeout 200 python bench.py --no-cpu-baseline --no-extras > "gpurun_out/bench_$name.json" 2> "gpurun_out/bench_$name.err"
  python - "$name" <<'PY'
import json, sys
try:
    d = json.load(open('gpurun_out/bench_%s.json' % sys.argv[1]))
    print(sys.argv[1], round(d['ms_per_step'], 3), d['stages']['ms'], round(d['e2e']['ms_per_step'], 2), d['parity_check']['identical'])
except Exception as e:
    print(sys.argv[1], 'bench line unreadable', e)
PY
}
for rep in 1 2; do
run_one default ""
for v in query_recommendation_system_b200/variants/*.so; do
  [ -e "$v" ] || continue
  run_one "$(basename "$v" .so)" "$PWD/$v"
done
done
echo "t=$((SECONDS-t0))"
