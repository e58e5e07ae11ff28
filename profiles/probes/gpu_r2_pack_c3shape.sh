#!/bin/bash
# C3 shape (100k-row shard): prepared rows in bank-class layout vs column order, stage times of the pipeline
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
t0=$SECONDS
for mode in class column class column; do
QRS_PREDICT_PACK=$mode QRS_CONFIG=C3 QRS_PROBE_N=100000 QRS_PROBE_REPS=5 timeout 200 python profiles/probes/pipeline_shape_probe.py 2>&1 | tail -1 | cut -c1-400
echo "$mode t=$((SECONDS-t0))"
done
