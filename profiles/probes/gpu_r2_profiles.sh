#!/bin/bash
# Round-2 final evidence in one call: scenario tests, launch list of the step, full captures of the predictor and the
# pair-scoring kernel (each after the same command exited 0 without ncu — the bench ran in the previous call and again here).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
t0=$SECONDS
timeout 200 python -m pytest tests/test_gpu_scenarios.py -x -q > gpurun_out/t_scen.log 2>&1
echo "scenario tests rc=$? t=$((SECONDS-t0))"; tail -15 gpurun_out/t_scen.log
timeout 120 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --no-extras > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err
echo "bench (short) rc=$? t=$((SECONDS-t0))"
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv \
  -k 'regex:^(minhash|perm_head|radix|scan_|pair|degree|fill_kernel|select_kernel|tile_index|seg_steps|pack_rows|single_rater|predict|topn|group_pad|route|unpack|widen|band)' \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity --no-extras > gpurun_out/ncu1.log 2>&1
echo "ncu launches rc=$? t=$((SECONDS-t0))"
timeout 120 ncu --set full --clock-control none --import-source on -k regex:predict_sparse -s 3 -c 1 -f \
  -o gpurun_out/predict python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --no-extras > gpurun_out/ncu2.log 2>&1
echo "ncu predict rc=$? t=$((SECONDS-t0))"
timeout 120 ncu --set full --clock-control none --import-source on -k regex:pairs_score -s 3 -c 1 -f \
  -o gpurun_out/pairs python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --no-extras > gpurun_out/ncu3.log 2>&1
echo "ncu pairs rc=$? t=$((SECONDS-t0))"
