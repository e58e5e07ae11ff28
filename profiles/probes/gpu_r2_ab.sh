#!/bin/bash
# A/B: MinHash tile scan with 16-byte vs 4-byte CSR loads (probe + bench stage), then the default bench line.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
t0=$SECONDS
rm -f gpurun_out/minhash_modes_C2.jsonl gpurun_out/minhash_modes_C3.jsonl
for rep in 1 2; do
QRS_CONFIG=C2 timeout 120 python profiles/probes/minhash_modes_probe.py 2>&1 | tail -1 | cut -c1-300
QRS_MINHASH_SCALAR_CSR=1 QRS_CONFIG=C2 timeout 120 python profiles/probes/minhash_modes_probe.py 2>&1 | tail -1 | cut -c1-300
done
QRS_CONFIG=C3 QRS_PROBE_N=200000 QRS_PROBE_REPS=5 timeout 200 python profiles/probes/minhash_modes_probe.py 2>&1 | tail -1 | cut -c1-300
QRS_MINHASH_SCALAR_CSR=1 QRS_CONFIG=C3 QRS_PROBE_N=200000 QRS_PROBE_REPS=5 timeout 200 python profiles/probes/minhash_modes_probe.py 2>&1 | tail -1 | cut -c1-300
echo "probes t=$((SECONDS-t0))"
timeout 240 python bench.py --no-cpu-baseline --no-extras > gpurun_out/bench_c2_v4.json 2> gpurun_out/bench.err
echo "bench (v4) rc=$? t=$((SECONDS-t0))"
QRS_MINHASH_SCALAR_CSR=1 timeout 240 python bench.py --no-cpu-baseline --no-extras > gpurun_out/bench_c2_scalar.json 2>> gpurun_out/bench.err
echo "bench (scalar) rc=$? t=$((SECONDS-t0))"
python - <<'PY'
import json
for f in ('v4', 'scalar'):
    d = json.load(open('gpurun_out/bench_c2_%s.json' % f))
    print(f, d['ms_per_step'], d['stages']['ms'], d['e2e']['ms_per_step'], d['parity_check']['identical'])
PY
