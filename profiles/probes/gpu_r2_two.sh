#!/bin/bash
# One `gpurun --gpus 2` call: banding tile-form tests + probe, hardware multi-GPU invariance (G = 2), weak C2 and strong C3 at N = 2.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=${N:-2}
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 "$@"; }
t0=$SECONDS
timeout 200 python -m pytest tests/test_gpu_stages.py -x -q -k "band or minhash_ragged or records" > gpurun_out/t_band.log 2>&1
echo "band tests rc=$? t=$((SECONDS-t0))"; tail -3 gpurun_out/t_band.log
for n in 100000 1000000; do
  timeout 60 python profiles/probes/band_probe.py $n 128 32 >> gpurun_out/band_probe.jsonl 2>> gpurun_out/band_probe.err
  QRS_BAND_PLAIN=1 timeout 60 python profiles/probes/band_probe.py $n 128 32 >> gpurun_out/band_probe.jsonl 2>> gpurun_out/band_probe.err
done
timeout 60 python profiles/probes/band_probe.py 1000000 256 64 >> gpurun_out/band_probe.jsonl 2>> gpurun_out/band_probe.err
QRS_BAND_PLAIN=1 timeout 60 python profiles/probes/band_probe.py 1000000 256 64 >> gpurun_out/band_probe.jsonl 2>> gpurun_out/band_probe.err
echo "band probe t=$((SECONDS-t0))"; cat gpurun_out/band_probe.jsonl | cut -c1-200
timeout 400 python -m pytest tests/test_gpu_multigpu.py -x -q > gpurun_out/t_multi.log 2>&1
echo "multi-GPU test rc=$? t=$((SECONDS-t0))"; tail -5 gpurun_out/t_multi.log
timeout 300 bash -c "$(declare -f run); N=$N run bench.py --gpus $N --no-extras --no-cpu-baseline" > gpurun_out/r2_c2_weak_g$N.json 2> gpurun_out/r2_c2_weak_g$N.err
echo "weak C2 rc=$? t=$((SECONDS-t0))"; tail -c 1200 gpurun_out/r2_c2_weak_g$N.json; tail -c 600 gpurun_out/r2_c2_weak_g$N.err
timeout 600 bash -c "$(declare -f run); N=$N run bench.py --gpus $N --config C3 --scaling strong --steps 3 --warmup 3 --no-extras --no-cpu-baseline" > gpurun_out/r2_c3_strong_g$N.json 2> gpurun_out/r2_c3_strong_g$N.err
echo "strong C3 rc=$? t=$((SECONDS-t0))"; tail -c 1200 gpurun_out/r2_c3_strong_g$N.json; tail -c 600 gpurun_out/r2_c3_strong_g$N.err
