#!/bin/bash
# C2 launch list (library kernels only) + the C3 line on one GPU.  Outputs -> gpurun_out/.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
t0=$SECONDS
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv \
  -k 'regex:^(minhash|perm_head|radix|scan_|pair|degree|fill_kernel|select_kernel|tile_index|seg_steps|pack_rows|single_rater|predict|topn|group_pad|route|unpack|widen|band)' \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity --no-extras > gpurun_out/ncu1.log 2>&1
echo "ncu launches rc=$? t=$((SECONDS-t0))"
timeout 1200 python bench.py --config C3 --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
echo "bench C3 rc=$? t=$((SECONDS-t0))"; tail -c 1500 gpurun_out/bench_c3.json; echo; tail -3 gpurun_out/bench_c3.err
