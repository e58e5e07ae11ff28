#!/bin/bash
# One 1-GPU call: MinHash tests (128-bit CSR loads), MinHash modes at a C4-shaped shard, e2e chunk schedules, copy probe.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
t0=$SECONDS
timeout 300 python -m pytest tests/test_gpu_stages.py -x -q -k "minhash" > gpurun_out/t_minhash.log 2>&1
echo "minhash tests rc=$? t=$((SECONDS-t0))"; tail -3 gpurun_out/t_minhash.log
rm -f gpurun_out/minhash_modes_C4.jsonl gpurun_out/minhash_modes_C2.jsonl
QRS_CONFIG=C2 timeout 120 python profiles/probes/minhash_modes_probe.py > gpurun_out/mh_c2.log 2>&1
echo "minhash C2 probe rc=$? t=$((SECONDS-t0))"; tail -1 gpurun_out/mh_c2.log | cut -c1-400
QRS_CONFIG=C4 QRS_PROBE_N=131072 QRS_PROBE_REPS=3 timeout 300 python profiles/probes/minhash_modes_probe.py > gpurun_out/mh_c4_a.log 2>&1
echo "minhash C4 probe rc=$? t=$((SECONDS-t0))"; tail -1 gpurun_out/mh_c4_a.log | cut -c1-400
QRS_CONFIG=C4 QRS_PROBE_N=131072 QRS_PROBE_REPS=3 QRS_PROBE_R=16384 timeout 300 python profiles/probes/minhash_modes_probe.py > gpurun_out/mh_c4_b.log 2>&1
echo "minhash C4 probe (R=16384) rc=$? t=$((SECONDS-t0))"; tail -1 gpurun_out/mh_c4_b.log | cut -c1-400
timeout 200 python profiles/probes/e2e_chunks_probe.py > gpurun_out/e2e_chunks.jsonl 2> gpurun_out/e2e_chunks.err
echo "e2e chunk probe rc=$? t=$((SECONDS-t0))"; cat gpurun_out/e2e_chunks.jsonl
timeout 60 python profiles/probes/copy_probe.py > gpurun_out/copy_probe_g1.json 2> gpurun_out/copy_probe_g1.err
echo "copy probe rc=$? t=$((SECONDS-t0))"; cat gpurun_out/copy_probe_g1.json
