#!/bin/bash
# One gpurun call: parity of the tile-scan MinHash kernel, its time against the per-set scan, the bench line,
# the launch list and one full ncu capture of the new kernel, then the whole GPU suite.  Outputs → gpurun_out/.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
t0=$SECONDS
timeout 240 python -m pytest tests/test_gpu_stages.py -m gpu -x -q -k minhash > gpurun_out/t_minhash.log 2>&1
echo "minhash tests rc=$? t=$((SECONDS-t0))"; tail -3 gpurun_out/t_minhash.log
timeout 120 python profiles/probes/minhash_modes_probe.py > gpurun_out/probe.log 2>&1
echo "probe rc=$? t=$((SECONDS-t0))"; tail -2 gpurun_out/probe.log
timeout 240 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench.err
echo "bench rc=$? t=$((SECONDS-t0))"; head -c 600 gpurun_out/bench_c2.json; tail -2 gpurun_out/bench.err
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:minhash -c 40 --csv \
  --log-file gpurun_out/minhash_launches.csv python profiles/probes/minhash_modes_probe.py > gpurun_out/ncu1.log 2>&1
echo "ncu launches rc=$? t=$((SECONDS-t0))"
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/t_all.log 2>&1
echo "all gpu tests rc=$? t=$((SECONDS-t0))"; tail -3 gpurun_out/t_all.log
timeout 150 ncu --set full --clock-control none --import-source on -k regex:minhash_tile -s 3 -c 1 -f \
  -o gpurun_out/minhash_tile python profiles/probes/minhash_modes_probe.py > gpurun_out/ncu2.log 2>&1
echo "ncu full rc=$? t=$((SECONDS-t0))"
