#!/bin/bash
# Second `gpurun --gpus 8` call: C4 (LSH stages) with the gather-form MinHash, weak C2 with the final kernels.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=${N:-8}
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 "$@"; }
t0=$SECONDS
timeout 600 bash -c "$(declare -f run); N=$N run bench.py --gpus $N --config C4 --scaling strong --lsh-only --steps 3 --warmup 3 --no-extras --no-cpu-baseline" > gpurun_out/r2_c4_lsh_g$N.json 2> gpurun_out/r2_c4_lsh_g$N.err
echo "C4 (LSH stages) rc=$? t=$((SECONDS-t0))"; tail -c 900 gpurun_out/r2_c4_lsh_g$N.json; echo; tail -c 400 gpurun_out/r2_c4_lsh_g$N.err
timeout 300 bash -c "$(declare -f run); N=$N run bench.py --gpus $N --no-extras --no-cpu-baseline" > gpurun_out/r2_c2_weak_g$N.json 2> gpurun_out/r2_c2_weak_g$N.err
echo "weak C2 rc=$? t=$((SECONDS-t0))"; tail -c 700 gpurun_out/r2_c2_weak_g$N.json; echo
