#!/bin/bash
# Round-end evidence in one gpurun call: the default bench line, then (after it exited 0 without ncu) the ncu
# launch list of the bench command and one full capture of the dominant kernel.  Outputs → gpurun_out/.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
t0=$SECONDS
timeout 60 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench.err
echo "bench rc=$? t=$((SECONDS-t0))"; head -c 300 gpurun_out/bench_c2.json; echo
# -k matches the kernel's BASE name (no namespace): `regex:qrs::` selects nothing with this ncu
timeout 75 ncu --metrics gpu__time_duration.sum --clock-control none -c 260 --csv --log-file gpurun_out/launches.csv \
  -k 'regex:^(minhash|perm_head|radix|scan_|pair|degree|fill_kernel|select_kernel|tile_index|single_rater|predict|topn)' \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu1.log 2>&1
echo "ncu launches rc=$? t=$((SECONDS-t0))"
timeout 60 ncu --set full --clock-control none --import-source on -k regex:predict_sparse_kernel -s 3 -c 1 -f \
  -o gpurun_out/predict python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu2.log 2>&1
echo "ncu full rc=$? t=$((SECONDS-t0))"
