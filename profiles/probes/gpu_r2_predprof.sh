#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
t0=$SECONDS
timeout 120 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --no-extras > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err
echo "bench (short) rc=$? t=$((SECONDS-t0))"
timeout 120 ncu --set full --clock-control none --import-source on -k regex:predict_sparse -s 3 -c 1 -f \
  -o gpurun_out/predict python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --no-extras > gpurun_out/ncu2.log 2>&1
echo "ncu predict rc=$? t=$((SECONDS-t0))"
