#!/bin/bash
# Round-2 state in one gpurun call: GPU suite, default bench line, ncu launch list, full capture of the predictor.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
t0=$SECONDS
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/t_all.log 2>&1
echo "all gpu tests rc=$? t=$((SECONDS-t0))"; tail -4 gpurun_out/t_all.log
timeout 240 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench.err
echo "bench rc=$? t=$((SECONDS-t0))"; head -c 600 gpurun_out/bench_c2.json; echo; tail -3 gpurun_out/bench.err
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity --no-extras > gpurun_out/ncu1.log 2>&1
echo "ncu launches rc=$? t=$((SECONDS-t0))"
timeout 90 ncu --set full --clock-control none --import-source on -k regex:predict_sparse -s 3 -c 1 -f \
  -o gpurun_out/predict python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --no-extras > gpurun_out/ncu2.log 2>&1
echo "ncu full rc=$? t=$((SECONDS-t0))"
