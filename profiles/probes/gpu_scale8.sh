#!/bin/bash
# One `gpurun --gpus 8` call: weak-scaled C2 at N = 8 (the SCALE shape) and C3 split over 8 ranks (strong).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=${N:-8}
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 "$@"; }
t0=$SECONDS
timeout 300 bash -c "$(declare -f run); N=$N run bench.py --gpus $N --no-extras --no-cpu-baseline" > gpurun_out/r2_c2_weak_g$N.json 2> gpurun_out/r2_c2_weak_g$N.err
echo "weak C2 rc=$? t=$((SECONDS-t0))"
if [ -n "$WITH_C3" ]; then
timeout 600 bash -c "$(declare -f run); N=$N run bench.py --gpus $N --config C3 --scaling strong --steps 3 --warmup 3 --no-extras --no-cpu-baseline" > gpurun_out/r2_c3_strong_g$N.json 2> gpurun_out/r2_c3_strong_g$N.err
echo "strong C3 rc=$? t=$((SECONDS-t0))"
fi
tail -c 600 gpurun_out/r2_c2_weak_g$N.err
