#!/bin/bash
# One `gpurun --gpus 8` call: weak C2 at N = 8 with remote neighbour rows staged (QRS_STAGE_PEER_ROWS=1) / read in place (=0).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=${N:-8}
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 "$@"; }
show() { python - <<PY
import json
for l in open('$1'):
    if l.startswith('{"metric'):
        d = json.loads(l)
        print(d['ms_per_step'], d['stages']['ms'], d['e2e']['ms_per_step'], d['parity_check']['identical'], (d.get('exchanges') or {}).get('ratings(peer reads)'))
PY
}
t0=$SECONDS
timeout 200 python -m pytest tests/test_gpu_predict_shapes.py -x -q -k "another_rank" > gpurun_out/t_stage.log 2>&1
echo "staging test rc=$? t=$((SECONDS-t0))"; tail -3 gpurun_out/t_stage.log
QRS_STAGE_PEER_ROWS=1 timeout 300 bash -c "$(declare -f run); N=$N run bench.py --gpus $N --no-extras --no-cpu-baseline" > gpurun_out/r2_c2_weak_g${N}_staged.json 2> gpurun_out/r2_c2_weak_g${N}_staged.err
echo "weak C2 (staged) rc=$? t=$((SECONDS-t0))"; show gpurun_out/r2_c2_weak_g${N}_staged.json; tail -c 300 gpurun_out/r2_c2_weak_g${N}_staged.err
QRS_STAGE_PEER_ROWS=0 timeout 300 bash -c "$(declare -f run); N=$N run bench.py --gpus $N --no-extras --no-cpu-baseline" > gpurun_out/r2_c2_weak_g${N}_inplace.json 2> gpurun_out/r2_c2_weak_g${N}_inplace.err
echo "weak C2 (in place) rc=$? t=$((SECONDS-t0))"; show gpurun_out/r2_c2_weak_g${N}_inplace.json; tail -c 300 gpurun_out/r2_c2_weak_g${N}_inplace.err
