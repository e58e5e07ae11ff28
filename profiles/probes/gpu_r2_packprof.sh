#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
t0=$SECONDS
timeout 200 ncu --set full --clock-control none --import-source on -k regex:pack_rows -s 3 -c 1 -f \
  -o gpurun_out/pack python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --no-extras > gpurun_out/ncu_pack.log 2>&1
echo "ncu pack rc=$? t=$((SECONDS-t0))"
