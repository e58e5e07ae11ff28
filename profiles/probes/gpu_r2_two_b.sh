#!/bin/bash
# One `gpurun --gpus 2` call: hardware multi-GPU invariance (G = 2, both row layouts), weak C2 at N = 2.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=${N:-2}
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 "$@"; }
t0=$SECONDS
timeout 400 python -m pytest tests/test_gpu_multigpu.py -x -q > gpurun_out/t_multi.log 2>&1
echo "multi-GPU test rc=$? t=$((SECONDS-t0))"; tail -5 gpurun_out/t_multi.log
QRS_PREDICT_PACK=class timeout 400 python -m pytest tests/test_gpu_multigpu.py -x -q > gpurun_out/t_multi_class.log 2>&1
echo "multi-GPU test (class layout forced) rc=$? t=$((SECONDS-t0))"; tail -5 gpurun_out/t_multi_class.log
timeout 300 bash -c "$(declare -f run); N=$N run bench.py --gpus $N --no-extras --no-cpu-baseline" > gpurun_out/r2_c2_weak_g$N.json 2> gpurun_out/r2_c2_weak_g$N.err
echo "weak C2 rc=$? t=$((SECONDS-t0))"; python - <<PY
import json
d = json.load(open('gpurun_out/r2_c2_weak_g$N.json'))
print(d['ms_per_step'], d['stages']['ms'], d['e2e']['ms_per_step'], d['parity_check'])
PY
tail -c 600 gpurun_out/r2_c2_weak_g$N.err
