#!/bin/bash
# One gpurun call: parity (minhash + predict tests) of the in-tree library, the MinHash mode probe for every
# variant library under query_recommendation_system_b200/variants/, the bench line, a full ncu capture of the
# tile kernel and the whole GPU suite.  Outputs → gpurun_out/.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
t0=$SECONDS
timeout 200 python -m pytest tests/test_gpu_stages.py -m gpu -x -q -k "minhash or predict" > gpurun_out/t_fast.log 2>&1
echo "minhash+predict tests rc=$? t=$((SECONDS-t0))"; tail -3 gpurun_out/t_fast.log
timeout 100 python profiles/probes/minhash_modes_probe.py > gpurun_out/probe_default.log 2>&1
echo "probe default rc=$? t=$((SECONDS-t0))"; tail -1 gpurun_out/probe_default.log
for v in query_recommendation_system_b200/variants/*.so; do
  [ -e "$v" ] || continue
  name=$(basename "$v" .so)
  QRS_B200_LIB="$PWD/$v" timeout 100 python profiles/probes/minhash_modes_probe.py > "gpurun_out/probe_$name.log" 2>&1
  echo "probe $name rc=$? t=$((SECONDS-t0))"; tail -1 "gpurun_out/probe_$name.log"
done
timeout 240 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench.err
echo "bench rc=$? t=$((SECONDS-t0))"; python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/bench_c2.json'))
    print(d['ms_per_step'], d['stages']['ms'], d['e2e']['ms_per_step'])
except Exception as e:
    print('bench line unreadable', e)
PY
tail -2 gpurun_out/bench.err
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/t_all.log 2>&1
echo "all gpu tests rc=$? t=$((SECONDS-t0))"; tail -3 gpurun_out/t_all.log
timeout 150 ncu --set full --clock-control none --import-source on -k regex:minhash_tile -s 3 -c 1 -f \
  -o gpurun_out/minhash_tile python profiles/probes/minhash_modes_probe.py > gpurun_out/ncu2.log 2>&1
echo "ncu full rc=$? t=$((SECONDS-t0))"
