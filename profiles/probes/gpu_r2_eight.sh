#!/bin/bash
# One `gpurun --gpus 8` call: hardware invariance for G = 2, 4, 8; weak C2, strong C3, C4 (LSH stages) at N = 8;
# the copy-only probe behind the 8-GPU e2e number.  Outputs -> gpurun_out/.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=${N:-8}
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 "$@"; }
t0=$SECONDS
timeout 500 python -m pytest tests/test_gpu_multigpu.py -x -q > gpurun_out/t_multi_g$N.log 2>&1
echo "multi-GPU tests rc=$? t=$((SECONDS-t0))"; tail -3 gpurun_out/t_multi_g$N.log
timeout 300 bash -c "$(declare -f run); N=$N run bench.py --gpus $N --no-extras --no-cpu-baseline" > gpurun_out/r2_c2_weak_g$N.json 2> gpurun_out/r2_c2_weak_g$N.err
echo "weak C2 rc=$? t=$((SECONDS-t0))"; tail -c 700 gpurun_out/r2_c2_weak_g$N.json; echo
timeout 400 bash -c "$(declare -f run); N=$N run bench.py --gpus $N --config C3 --scaling strong --steps 3 --warmup 3 --no-extras --no-cpu-baseline" > gpurun_out/r2_c3_strong_g$N.json 2> gpurun_out/r2_c3_strong_g$N.err
echo "strong C3 rc=$? t=$((SECONDS-t0))"; tail -c 700 gpurun_out/r2_c3_strong_g$N.json; echo
timeout 90 bash -c "$(declare -f run); N=$N run profiles/probes/copy_probe.py" > gpurun_out/r2_copy_probe_g$N.json 2> gpurun_out/r2_copy_probe_g$N.err
echo "copy probe rc=$? t=$((SECONDS-t0))"; cat gpurun_out/r2_copy_probe_g$N.json
timeout 600 bash -c "$(declare -f run); N=$N run bench.py --gpus $N --config C4 --scaling strong --lsh-only --steps 2 --warmup 3 --no-extras --no-cpu-baseline" > gpurun_out/r2_c4_lsh_g$N.json 2> gpurun_out/r2_c4_lsh_g$N.err
echo "C4 (LSH stages) rc=$? t=$((SECONDS-t0))"; tail -c 900 gpurun_out/r2_c4_lsh_g$N.json; echo; tail -c 800 gpurun_out/r2_c4_lsh_g$N.err
