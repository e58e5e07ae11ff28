#!/bin/bash
# ncu launch list of the bench command (the library's kernels by base name), bounded for a short gpurun call.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 44 ncu --metrics gpu__time_duration.sum --clock-control none -c 220 --csv --log-file gpurun_out/launches.csv \
  -k 'regex:^(minhash|perm_head|radix|scan_|pair|degree|fill_kernel|select_kernel|tile_index|single_rater|predict|topn)' \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu1.log 2>&1
echo "ncu launches rc=$?"; wc -l gpurun_out/launches.csv
