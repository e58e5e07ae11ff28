#!/bin/bash
# Final round-2 evidence in one short call: the whole GPU suite, the default bench line, the launch list of the step.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
t0=$SECONDS
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/t_all.log 2>&1
echo "all gpu tests rc=$? t=$((SECONDS-t0))"; tail -4 gpurun_out/t_all.log
timeout 120 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench.err
echo "bench rc=$? t=$((SECONDS-t0))"; python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/bench_c2.json'))
    print(d['ms_per_step'], d['stages']['ms'], d['e2e']['ms_per_step'], d['parity_check']['identical'], d['roofline']['frac'])
except Exception as e:
    print('bench line unreadable', e)
PY
tail -2 gpurun_out/bench.err
timeout 45 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv \
  -k 'regex:^(minhash|perm_head|radix|scan_|pair|degree|fill_kernel|select_kernel|tile_index|seg_steps|pack_rows|single_rater|predict|topn|group_pad|route|unpack|widen|band)' \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity --no-extras > gpurun_out/ncu1.log 2>&1
echo "ncu launches rc=$? t=$((SECONDS-t0))"
