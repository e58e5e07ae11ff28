#!/usr/bin/env python
"""Benchmark of the LSH → top-k CF hot path (BASELINE.json metric: MinHash sigs/s, LSH candidate
pairs/s, predicted ratings/s) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config C2] [--impl ours|reference]

A "step" is one pass of the whole path (signatures → bucket records → grouping → candidate pairs +
cosine → neighbour tables → predictions) over one synthetic utility matrix of the named config
(generator.py schema, planted taste groups).  `value` = sets signed AND carried through
the whole path per second, inputs resident in HBM; `stages` breaks it into sigs/s, pairs/s and
ratings/s.  `e2e` is the same pass through the host-facing call with pinned HOST buffers (H2D of the
CSR and D2H of the predictions inside the timed region).  N > 1: one rank per GPU (torchrun), row
shards of the config's size per GPU (weak scaling), NCCL exchanges inside the step.

`--impl reference` times the UNMODIFIED reference (oracle/refarm.py drives the modules build() staged into
oracle/_ref; kind = "reference" — the loop-level port oracle/port.py, kind = "port", only if nothing was staged)
on a bounded closed sample of the same workload; the `cpu_baseline` leg of the GPU arm runs the same code on
the same sample size, and the GPU pipeline is run on that very sample next to it.

After the timed loop (outside it) the last step's outputs are compared with the CPU oracle at the benchmarked
shape (`parity_check`; N > 1: with a single-GPU run of the same matrix at a reduced n) — a mismatch exits 1.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "MinHash sigs/s through the full LSH->CF path (signatures, banding, candidate pairs, cosine, top-k, predictions)"
UNIT = "sigs/s"
DTYPE = "int32/int64 (signatures, buckets, cosine dot) + f64 (similarity, prediction)"
DATA_SEED = 20260102          # 20260101 + config id
PERM_SEED = 7


def perm_table(m, K, seed=PERM_SEED):
    """K successive np.random.permutation(m) from the legacy global RNG (recommender.py:118-120)."""
    np.random.seed(seed)
    tab = np.empty((m, K), dtype=np.int32)
    for i in range(K):
        tab[:, i] = np.random.permutation(m)
    return tab


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region: the sampler runs from before the
    warm-up (nvidia-smi needs a moment to start), every line carries a timestamp, and only the lines that
    fall between mark_start() and mark_end() are summarised (all lines under load if the region was
    shorter than the sampling period)."""
    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None
        self.t0 = self.t1 = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def mark_start(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

        def summarise(lines):
            sm, mx, reasons = [], [], set()
            for _, ln in lines:
                f = [x.strip() for x in ln.split(",")]
                if len(f) < 10:
                    continue
                try:
                    sm.append(float(f[2]))
                    mx.append(float(f[3]))
                except ValueError:
                    continue
                for nm, v in zip(names, f[6:10]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            return sm, mx, reasons

        inside = [x for x in self.lines if self.t0 is not None and self.t0 <= x[0] <= (self.t1 or x[0]) + 0.03]
        window = "timed region"
        sm, mx, reasons = summarise(inside)
        if not sm:
            window = "warm-up + timed region (the timed region was shorter than one sampling period)"
            sm, mx, reasons = summarise(self.lines)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "window": window, "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------ CPU side
def cpu_sample(config, n_sets):
    from query_recommendation_system_b200 import synth
    n, m, density, K, b, k = synth.config_shape(config)
    rowptr, colidx, val = synth.planted_csr(n_sets, m, density, seed=DATA_SEED, device="cpu")
    return (rowptr.numpy(), colidx.numpy(), val.numpy()), (m, K, b, k)


def cpu_sample_sets(config, steps, warmup):
    """Sets of the closed CPU sample: the same number in both arms (the driver passes both the same --steps /
    --warmup), sized so that (steps + warmup) passes of the reference stay within a few minutes."""
    per_set_s = {"C2": 0.075, "C3": 0.9, "C4": 9.0, "tiny": 0.004}.get(config, 0.1)   # measured: ~m * 7.5 us per set
    budget_s = 170.0
    n = int(budget_s / (per_set_s * max(1, steps + warmup)))
    return int(max(8, min(64, n // 8 * 8)))


def run_cpu_path(config, n_sets, want_results=False):
    """One pass of the CPU path over the closed sample: the unmodified reference when oracle/_ref is staged
    (kind "reference"), else the loop-level port (kind "port")."""
    from oracle import baseline, refarm
    (rowptr, colidx, val), (m, K, b, k) = cpu_sample(config, n_sets)
    if refarm.available():
        t, counts, res = refarm.run_reference_pipeline(rowptr, colidx, val, m, K, b, k, perm_seed=PERM_SEED)
        kind = "reference"
    else:
        t, counts = baseline.run_port_pipeline(rowptr, colidx, val, m, K, b, k, perm_seed=PERM_SEED)
        res, kind = None, "port"
    return (t, counts, kind, res) if want_results else (t, counts, kind)


def cpu_sample_note(config, n_sets, kind):
    how = ("the UNMODIFIED reference (oracle/_ref: Recommender.compute_signatures via the compute_shingles seam, lsh.LSH, "
           "the cosine/top-k block with sklearn, Recommender.compute_scores' numba cell loop)" if kind == "reference"
           else "oracle/port.py (loop-level port; oracle/_ref was not staged)")
    return ("first %d sets of the %s matrix as a closed sample (LSH and CF over these sets only), whole path per step; %s; "
            "1 Python thread, like the reference. The reference's per-set cost falls with the sample size (every "
            "permutation walks all m elements whatever n is), so this sample understates its throughput at the full n: "
            "no same-size ratio is claimed" % (n_sets, config, how))


def reference_arm(args):
    """bench.py --impl reference: the reference's CPU path on a bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import refarm
    steps, warmup = args.steps, args.warmup
    n_sets = args.cpu_sets or cpu_sample_sets(args.config, steps, warmup)
    if refarm.available():
        refarm.warm_numba()
    for _ in range(warmup):
        run_cpu_path(args.config, n_sets)
    t0 = time.perf_counter()
    stage = {}
    for _ in range(steps):
        t, counts, kind = run_cpu_path(args.config, n_sets)
        for kx, v in t.items():
            stage[kx] = stage.get(kx, 0.0) + v
    wall = time.perf_counter() - t0
    path_s = stage["total"] / steps
    value = n_sets / path_s
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": path_s * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
        "config": workload_config(args.config, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": kind, "sample": cpu_sample_note(args.config, n_sets, kind),
                         "sample_sets": n_sets, "host_cores": os.cpu_count(),
                         "stage_seconds_per_step": {kx: v / steps for kx, v in stage.items()}, "counts": counts},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": wall,
    }
    print(json.dumps(line))


def onchip_roofline(wavefronts, ms, sm_count, sm_mhz):
    """Second roof of a shared-memory-bound kernel: `wavefronts` shared-memory wavefronts per launch (ncu:
    l1tex__data_pipe_lsu_wavefronts_mem_shared.sum) at one wavefront per SM per clock, against the measured time."""
    floor_ms = wavefronts / (sm_count * sm_mhz * 1e6) * 1e3
    return {"bound": "shared-memory data pipe", "achieved": wavefronts / (ms * 1e-3) / 1e9, "peak": sm_count * sm_mhz * 1e-3,
            "unit": "Gwavefront/s", "frac": floor_ms / ms, "wavefronts_per_launch": int(wavefronts), "floor_ms": floor_ms}


def workload_config(config, n_gpus):
    """Identical in both arms (same keys, same values) for the same --config / --gpus."""
    from query_recommendation_system_b200 import synth
    n, m, density, K, b, k = synth.config_shape(config)
    csr_mb = (n * m * density * 5 + 8 * n) / 1e6
    return {"workload": "%s: %d users x %d queries, %.3g%% density, K=%d hashes, b=%d bands x r=%d rows, top-k=%d, "
                        "planted taste groups of 32 (SURVEY.md 8d), ratings 1..100"
                        % (config, n, m, density * 100, K, b, K // b, k),
            "sets_per_gpu": n, "elements": m, "density": density, "K": K, "b": b, "r": K // b, "k": k,
            "parallelism": "row shards x%d, bucket all-to-all" % n_gpus if n_gpus > 1 else "single GPU",
            "binding": "B (set = user row of the utility matrix, element = rated query)",
            "permutations": "np.random.seed(%d); K x np.random.permutation(m) (setup, not timed)" % PERM_SEED,
            "l2": "GPU arm: inputs (~%.0f MB CSR per GPU) exceed the 126 MB L2, no flush; CPU arm: not applicable" % csr_mb}


def sweep(args):
    """BASELINE.json configs[4]: band/row sweep (b x r in {16x8, 32x4, 64x2}, K = 128) on the named config's matrix —
    candidate-pair volume vs throughput of the grouping + pair/similarity kernels.  One JSON line, not the bench line."""
    import torch
    from query_recommendation_system_b200 import stages, synth
    from query_recommendation_system_b200.lsh import group_and_pair
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    n, m, density, _, _, _ = synth.config_shape(args.config)
    K = 128
    rowptr, colidx, _ = synth.planted_csr(n, m, density, seed=DATA_SEED, device=dev)
    tab = torch.from_numpy(perm_table(m, K)).to(dev)
    head = stages.perm_head(tab, stages.choose_head_ranks(m, colidx.numel() / n))
    sig, norm2, _ = stages.minhash(rowptr, colidx, m, tab, head, b=0)
    rows = []
    for b in (16, 32, 64):
        times = []
        for it in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            records, _ = stages.band_records(sig, b, 0, want_norm2=False)
            pairs, sims, n_slots = group_and_pair(records, n, b, sig, norm2, b, True)
            e1.record()
            torch.cuda.synchronize()
            if it > 0:
                times.append(e0.elapsed_time(e1))
        ms = float(np.median(times))
        rows.append({"b": b, "r": K // b, "pair_slots": int(n_slots), "candidate_pairs": int(pairs.numel()), "ms": ms,
                     "candidate_pairs_per_s": pairs.numel() / (ms * 1e-3), "pair_slots_per_s": n_slots / (ms * 1e-3)})
    print(json.dumps({"sweep": "banding + grouping + pairs + cosine, K=128", "config": args.config, "sets": n,
                      "elements": m, "density": density, "rows": rows}))


# ------------------------------------------------------------------------------------------ parity (outside the timed loop)
def decode_user(rs, rc, cols, preds, cols_mode, tile, t):
    """(columns, predictions) of local user t from the sparse prediction outputs (host arrays)."""
    nt = rs.shape[1]
    ps = np.concatenate([preds[rs[t, x]:rs[t, x] + rc[t, x]] for x in range(nt)])
    if cols_mode == 2:
        words = tile // 32
        bm = cols.view(np.uint32).reshape(rs.shape[0], nt, words)[t]
        bits = np.unpackbits(bm.view(np.uint8).reshape(nt, words * 4), axis=1, bitorder="little")
        cs = np.concatenate([x * tile + np.flatnonzero(bits[x]) for x in range(nt)])
        return cs[ps != 0].astype(np.int64), ps[ps != 0].astype(np.int64)
    c = cols.view(np.uint16) if cols_mode == 1 else cols
    cs = np.concatenate([c[rs[t, x]:rs[t, x] + rc[t, x]] for x in range(nt)]).astype(np.int64)
    return cs, ps.astype(np.int64)


def parity_check_single(config, out, rowptr, colidx, val, tab, m, K, b, k, sig_rows=4096, pair_sample=100_000, users=256):
    """The last step's outputs against the CPU oracle (oracle/spec.py) at the benchmarked shape:
    signatures of `sig_rows` sampled rows, the FULL candidate set, cosines of `pair_sample` pairs, the FULL
    neighbour tables, the predictions of `users` sampled users.  Bit-exact everywhere."""
    import torch
    from oracle import spec
    t0 = time.perf_counter()
    rng = np.random.default_rng(12345)
    n = rowptr.numel() - 1
    rp, ci, va = rowptr.cpu().numpy(), colidx.cpu().numpy(), val.cpu().numpy()
    sig = out.sig.cpu().numpy().astype(np.int64)
    checks = {}
    rows = np.sort(rng.choice(n, min(sig_rows, n), replace=False))
    lens = rp[rows + 1] - rp[rows]
    sub_rp = np.zeros(rows.size + 1, np.int64)
    np.cumsum(lens, out=sub_rp[1:])
    sub_ci = np.concatenate([ci[rp[r]:rp[r + 1]] for r in rows]) if rows.size else np.zeros(0, np.int32)
    checks["signatures_rows_%d" % rows.size] = bool(np.array_equal(spec.signatures(sub_rp, sub_ci, tab), sig[rows]))
    n2 = out.norm2.cpu().numpy()
    checks["norm2"] = bool(np.array_equal(n2, (sig * sig).sum(axis=1)))
    cand = spec.candidates_packed_fast(sig, b)
    got = out.pairs.cpu().numpy().view(np.uint64)
    order = np.argsort(got, kind="stable")
    checks["candidate_set_full_%d" % cand.size] = bool(np.array_equal(got[order], cand))
    sims_got = out.sims.cpu().numpy()[order]
    pick = np.sort(rng.choice(cand.size, min(pair_sample, cand.size), replace=False)) if cand.size else np.zeros(0, np.int64)
    pr_all = np.stack([(cand >> np.uint64(32)).astype(np.int64), (cand & np.uint64(0xFFFFFFFF)).astype(np.int64)], 1)
    ok_cos = bool(checks["candidate_set_full_%d" % cand.size] and
                  np.array_equal(spec.cosine_milli(sig, pr_all[pick]), sims_got[pick]))
    checks["cosine_pairs_%d" % pick.size] = ok_cos
    tables_ok = False
    idx = v = cnt = None
    if checks["candidate_set_full_%d" % cand.size]:
        idx, v, cnt = spec.topk_tables(n, pr_all, sims_got / 1000.0, k)      # cosines verified on the sample above
        tables_ok = bool(np.array_equal(out.nbr_idx.cpu().numpy(), idx) and np.array_equal(out.nbr_val.cpu().numpy(), v)
                         and np.array_equal(out.nbr_cnt.cpu().numpy(), cnt))
    checks["neighbour_tables_full"] = tables_ok
    pred_ok = False
    n_cells = 0
    if tables_ok and out.pred_chunks:
        pred_ok = True
        urows = np.sort(rng.choice(n, min(users, n), replace=False))
        ref = spec.predict_user_branch_csr(rp, ci, va, m, idx, v, cnt, rows=urows.tolist())
        for ch in out.pred_chunks:
            lo, hi = ch["user_lo"], ch["user_lo"] + ch["n_users"]
            mine = urows[(urows >= lo) & (urows < hi)]
            if mine.size == 0:
                continue
            rs, rc = ch["row_start"].cpu().numpy(), ch["row_cnt"].cpu().numpy()
            cols, preds = ch["cols"].cpu().numpy(), ch["preds"].cpu().numpy()
            for u in mine.tolist():
                gc, gp = decode_user(rs, rc, cols, preds, ch["cols_mode"], ch["tile"], u - lo)
                n_cells += gc.size
                if not (np.array_equal(gc, ref[u][0]) and np.array_equal(gp, ref[u][1])):
                    pred_ok = False
    checks["predictions_users_%d" % min(users, n)] = pred_ok
    return {"config": config, "against": "oracle/spec.py (NumPy restatement pinned to the reference's goldens)",
            "identical": all(checks.values()), "checks": checks, "prediction_cells_compared": n_cells,
            "seconds": time.perf_counter() - t0}


# ------------------------------------------------------------------------------------------ GPU side
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default="C2")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sets", type=int, default=0, help="sets of the closed CPU sample (0: from --steps/--warmup)")
    ap.add_argument("--no-parity", action="store_true", help="skip the post-run parity check")
    ap.add_argument("--no-extras", action="store_true", help="skip the f-2/f-3/f-4 extras measured outside the step")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N > 1: 'weak' = one shard of the config's size per GPU (default), 'strong' = the config's n split over the GPUs")
    ap.add_argument("--sweep", action="store_true", help="band/row sweep of the LSH stages (extra, not the bench line)")
    ap.add_argument("--lsh-only", action="store_true",
                    help="stop after the neighbour tables (signatures .. top-k): for shapes whose predictions cannot be "
                         "materialised (C4: ~5e11 predictions); no e2e leg")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
        return
    if args.sweep:
        sweep(args)
        return

    import torch
    import torch.distributed as dist
    from query_recommendation_system_b200 import _lib, synth
    from query_recommendation_system_b200.pipeline import HostRunner, LshCfPipeline

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus or world == 1 and args.gpus == 1, "launch with torchrun for --gpus > 1"

    n_cfg, m, density, K, b, k = synth.config_shape(args.config)
    strong = args.scaling == "strong" and world > 1
    n = n_cfg // world if strong else n_cfg                 # rows per GPU
    lib = _lib.load()
    rowptr, colidx, val = synth.planted_csr(n, m, density, seed=DATA_SEED, row_base=rank * n, device=dev)
    nnz = int(colidx.numel())
    tab = perm_table(m, K)
    pipe = LshCfPipeline(m=m, K=K, b=b, k=k, perm_tab=tab, device=dev, profile=True)
    # the device-timed step writes the same compact column format the host-facing call ships
    pipe.cols_mode = 1 if m <= 65536 else 0

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    warm = max(args.warmup, 3)
    do_predict = not args.lsh_only
    for _ in range(warm):
        out = pipe.run(rowptr, colidx, val, predict=do_predict)
        if pipe.cols_mode == 1 and out.n_predictions * 16 > n * m:
            pipe.cols_mode = 2          # touched-column bitmap beats 2 bytes per prediction (HostRunner's "auto" rule)
    barrier()

    launches0 = lib.qrs_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stage_ms = {}
    barrier()
    sampler.mark_start()
    ev0.record()
    outs = []
    for _ in range(args.steps):
        out = pipe.run(rowptr, colidx, val, predict=do_predict)
        outs.append(out.stage_ms)
    ev1.record()
    barrier()
    sampler.mark_end()
    clocks = sampler.stop()
    launches = lib.qrs_launch_count() - launches0
    elapsed_ms = ev0.elapsed_time(ev1)
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    for d in outs:
        for kx, v in d.items():
            stage_ms[kx] = stage_ms.get(kx, 0.0) + v / args.steps
    if world > 1:       # per-stage maximum over the ranks (the step time is the max over ranks too)
        names = sorted(stage_ms)
        tt = torch.tensor([stage_ms[x] for x in names], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        stage_ms = dict(zip(names, [float(x) for x in tt.tolist()]))

    # per-rank counts -> whole-job totals
    counts = torch.tensor([out.pairs.numel(), out.n_slots, out.n_predictions, out.covered_cells], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(counts)
    n_pairs, n_slots, n_pred, covered = [int(x) for x in counts.tolist()]
    n_total = n * world
    ms_per_step = elapsed_ms / args.steps
    value = n_total / (ms_per_step * 1e-3)

    # algorithmic bytes of the two heavy kernels from the actual data
    lens = (rowptr[1:] - rowptr[:-1])
    sig_bytes = 4 * nnz + 8 * (n + 1) + 4 * n * K + 8 * n * b + 8 * n
    gathered = int(out.gathered_entries) if getattr(out, "gathered_entries", None) is not None else None
    if gathered is None:
        if world == 1:
            idx = out.nbr_idx.long().clamp(min=0)
            mask = torch.arange(k, device=dev)[None, :] < out.nbr_cnt[:, None]
            gathered = int((lens[idx] * mask).sum().item())
        else:
            gathered = int(out.nbr_cnt.sum().item()) * (nnz // n)
    pred_bytes = 5 * gathered + 5 * out.n_predictions
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    ms_sig = stage_ms.get("signatures+banding", float("nan"))
    ms_pred = stage_ms.get("predict", float("nan"))
    rl = {
        "signatures+banding": {"bound": "hbm", "achieved": sig_bytes / (ms_sig * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                               "bytes_per_launch": sig_bytes, "ms": ms_sig, "launches_per_step": 1},
        "predict": {"bound": "hbm", "achieved": pred_bytes / (ms_pred * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                    "bytes_per_launch": pred_bytes, "ms": ms_pred,
                    "note": "stage = predict kernel + tie fix-up + the step's one counter readback (the once-per-step row preparation is the "
                            "separate stage predict:prepare)"},
    }
    if not do_predict:
        del rl["predict"]
    for v in rl.values():
        v["frac"] = v["achieved"] / peak
    dominant = max(rl, key=lambda s: stage_ms.get(s, 0.0))
    roofline = dict(rl[dominant])
    kernel_name = "predict_sparse_kernel" if dominant == "predict" else "minhash_tile_kernel"
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")     # DRAM bytes per launch from the committed ncu capture
    tinfo = json.load(open(tpath)) if os.path.exists(tpath) else {}
    if world == 1 and args.config == "C2":
        traffic = tinfo.get(kernel_name)
    roofline.update({"kernel": kernel_name, "stage": dominant, "peak_source": peak_src + " (of measured)",
                     "traffic": traffic,
                     "traffic_source": "%s (ncu --set full capture of this command; not re-measured by this run)"
                                       % tinfo.get("source", "profiles/traffic.json")})
    try:    # the on-chip roof the dominant kernel is actually close to (counter from the same committed capture)
        wf = tinfo.get(kernel_name + ":smem_wavefronts") if traffic is not None else None
        if wf:
            roofline["onchip"] = onchip_roofline(float(wf), roofline["ms"], torch.cuda.get_device_properties(dev).multi_processor_count,
                                                 float(clocks.get("sm_mhz") or 1965.0))
    except Exception:
        pass

    # ---------------------------------------------------------------- parity of the step just timed (outside the timed region)
    parity = None
    if not args.no_parity:
        if world == 1:
            if n * K <= 300e6:
                parity = parity_check_single(args.config, out, rowptr, colidx, val, tab, m, K, b, k)
            else:
                parity = {"config": args.config, "skipped": "full-size oracle pass too large for the bench run"}
        else:
            from query_recommendation_system_b200 import selfcheck
            n_red = min(n, 16384 if m <= 200_000 else 4096)
            parity = selfcheck.multi_rank_invariance(n_red, m, density, K, b, k, tab, dev, seed=DATA_SEED)
            if rank == 0:
                parity = dict(parity, config="%s at %d sets per rank" % (args.config, n_red),
                              against="single-GPU run of the same %d x %d matrix on rank 0" % (n_red * world, m))

    extras = {}
    if world == 1 and not args.no_extras:
        extras = measure_extras(args, pipe, rowptr, colidx, val, n, m, nnz, peak, dev)

    exchanges = None
    if world > 1:
        exchanges = exchange_report(outs, out, stage_ms, dev, world)

    e2e_ms = e2e_wall_ms = None
    runner = None
    if do_predict:
        runner = HostRunner(pipe)
        # host inputs: the CSR of this rank's shard in pinned memory; column ids in the compact 16-bit wire format when the
        # matrix has at most 65536 queries (HostRunner widens them on the device, inside the timed call)
        h_rowptr, h_val = rowptr.cpu().pin_memory(), val.cpu().pin_memory()
        h_colidx = colidx.cpu()
        if m <= 65536:
            h_colidx = torch.from_numpy(h_colidx.numpy().astype(np.uint16))
        h_colidx = h_colidx.pin_memory()
        for _ in range(4):          # capacities, the column format and every pinned buffer settle within three calls
            res = runner(h_rowptr, h_colidx, h_val)
        barrier()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        e2e_steps = max(3, min(args.steps, 10))
        for _ in range(e2e_steps):
            res = runner(h_rowptr, h_colidx, h_val)
        e1.record()
        barrier()
        e2e_wall_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
        te = torch.tensor([e0.elapsed_time(e1) / e2e_steps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_ms = float(te.item())

    rc = 0
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
            "dtype": DTYPE, "data": "synthetic",
            "config": workload_config(args.config, world),
            "gpu_launches": int(launches), "clocks": clocks,
            "notes": {"capacity": "output capacities are learnt in the warm-up (a cold call sizes them with one counting pass of the "
                                  "predictor) and reused by the timed steps",
                      "columns": "timed step and e2e both write the compact column format (cols_mode %d)" % pipe.cols_mode},
        }
        line.update(extras)
        if world == 1 and not args.no_cpu_baseline:
            from oracle import refarm
            n_cpu = args.cpu_sets or cpu_sample_sets(args.config, args.steps, args.warmup)
            if refarm.available():
                refarm.warm_numba()
            tc, cc, kind, cres = run_cpu_path(args.config, n_cpu, want_results=True)
            line["cpu_baseline"] = {
                "value": n_cpu / tc["total"], "unit": UNIT, "cores": 1, "kind": kind, "sample_sets": n_cpu,
                "host_cores": os.cpu_count(), "seconds": tc, "counts": cc, "sample": cpu_sample_note(args.config, n_cpu, kind),
                "gpu_same_sample": gpu_on_cpu_sample(args.config, n_cpu, tab, dev, cres)}
        line["rooflines"] = rl
        line["roofline"] = roofline
        if e2e_ms is None:
            line["e2e"] = None
            line["notes"]["lsh_only"] = ("--lsh-only: the step ends with the neighbour tables; the predictions of this shape cannot be "
                                         "materialised, so there is no prediction stage and no e2e leg")
        else:
            line["e2e"] = {
                "value": n_total / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms, "wall_ms_per_step": e2e_wall_ms,
                "h2d_bytes_per_step": runner.h2d_bytes, "d2h_bytes_per_step": runner.d2h_bytes, "steps": e2e_steps,
                "call": "pipeline.HostRunner(pipe)(rowptr, colidx, val) with pinned host tensors (column ids uint16 when "
                        "m <= 65536); returns pinned sparse predictions + neighbour tables (ids, 1000 x similarity, counts)"}
        if exchanges is not None:
            line["exchanges"] = exchanges
        line["parity_check"] = parity
        # last in the line, so that a truncated tail still shows the per-stage split
        line["stages"] = {
            "sigs_per_s": n_total / (ms_sig * 1e-3),
            "candidate_pairs_per_s": n_pairs / (sum(stage_ms.get(s, 0.0) for s in ("grouping", "pairs+cosine")) * 1e-3),
            "pair_slots_per_s": n_slots / (sum(stage_ms.get(s, 0.0) for s in ("grouping", "pairs+cosine")) * 1e-3),
            "predicted_ratings_per_s": covered / (ms_pred * 1e-3) if do_predict else None,
            "nonzero_predictions_per_s": n_pred / (ms_pred * 1e-3) if do_predict else None,
            "counts": {"sets": n_total, "nnz_per_gpu": nnz, "candidate_pairs": n_pairs, "pair_slots": n_slots,
                       "nonzero_predictions": n_pred, "covered_cells": covered},
            "ms": {kx: round(v, 4) for kx, v in stage_ms.items()},
        }
        print(json.dumps(line))
        if parity is not None and parity.get("identical") is False:
            sys.stderr.write("PARITY CHECK FAILED: %s\n" % json.dumps(parity))
            rc = 1
    if world > 1:
        dist.destroy_process_group()
    if rc:
        sys.exit(rc)


def gpu_on_cpu_sample(config, n_sets, tab, dev, cpu_results):
    """The GPU pipeline on the very sample the CPU arm ran (identical inputs), its time, and — when the CPU arm was
    the unmodified reference — whether signatures, candidate set and predictions agree with the reference's own."""
    import torch
    from query_recommendation_system_b200.pipeline import LshCfPipeline
    (rowptr, colidx, val), (m, K, b, k) = cpu_sample(config, n_sets)
    pipe = LshCfPipeline(m=m, K=K, b=b, k=k, perm_tab=tab, device=dev, distributed=False)
    d = [torch.from_numpy(x).to(dev) for x in (rowptr, colidx, val)]
    for _ in range(3):
        o = pipe.run(*d)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        o = pipe.run(*d)
    e1.record()
    torch.cuda.synchronize()
    res = {"ms_per_step": e0.elapsed_time(e1) / 5, "sets": n_sets,
           "note": "launch-latency bound at this size (about 30 kernel launches); shown so that CPU and GPU are seen on identical inputs"}
    if cpu_results is not None:
        sig_ok = bool(np.array_equal(o.sig.cpu().numpy().astype(np.int64), cpu_results["sig"]))
        got = o.pairs.cpu().numpy().view(np.uint64)
        want = np.array(sorted((int(i) << 32) | int(j) for i, j in cpu_results["cand"]), dtype=np.uint64)
        cand_ok = bool(np.array_equal(np.sort(got), want))
        rs, rc = o.pred_row_start.cpu().numpy(), o.pred_row_cnt.cpu().numpy()
        cols, preds = o.pred_cols.cpu().numpy(), o.pred_vals.cpu().numpy()
        final = cpu_results["final"]
        differing = 0
        for u in range(n_sets):
            gc, gp = decode_user(rs, rc, cols, preds, 0, 0, u)
            dense = np.zeros(m, dtype=np.int64)
            dense[gc] = gp
            lo, hi = rowptr[u], rowptr[u + 1]
            dense[colidx[lo:hi]] = np.where(val[lo:hi] != 0, val[lo:hi], dense[colidx[lo:hi]])
            differing += int((dense != final[u]).sum())
        res["vs_reference_outputs"] = {"signatures_identical": sig_ok, "candidate_set_identical": cand_ok,
                                       "prediction_cells": int(n_sets * m), "prediction_cells_differing": differing,
                                       "note": "cells may differ only where the reference's unspecified neighbour order "
                                               "(set iteration + unstable argsort, recommender.py:185-210) changes an fp64 sum at an exact .5"}
    return res


def exchange_report(outs, out, stage_ms, dev, world):
    """Bytes each exchange moved per rank (sent over NVLink) and the fraction of the measured peer-copy rate
    (770 GB/s per direction per GPU, B200_PROFILING.md; 900 nominal) its stage time corresponds to."""
    import torch
    import torch.distributed as dist
    ex = getattr(out, "exchange_bytes", None) or {}
    names = sorted(ex)
    if not names:
        return None
    tt = torch.tensor([float(ex[x]) for x in names], dtype=torch.float64, device=dev)
    mx = tt.clone()
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    rep = {}
    for name, bts in zip(names, mx.tolist()):
        ms = stage_ms.get("exchange:ratings" if name.startswith("ratings") else "exchange:" + name)
        rep[name] = {"bytes_sent_per_rank_max": int(bts), "ms": ms,
                     "nvlink_frac_of_770GBps": (bts / (ms * 1e-3) / 770e9) if ms else None}
        if name == "signatures":
            # all-gather on a second stream and communicator, behind the record exchange and the grouping: `ms` is its own
            # duration, `exposed_ms` what the main stream waited for it
            bg = stage_ms.get("background:signatures")
            rep[name].update({"ms": bg, "exposed_ms": ms, "nvlink_frac_of_770GBps": (bts / (bg * 1e-3) / 770e9) if bg else None})
        if name.startswith("ratings"):
            staged = bool(getattr(out, "peer_rows_staged", False))
            rep[name]["staged"] = staged
            rep[name]["note"] = ("neighbour rows of other ranks copied into a local buffer before the prediction kernel "
                                 "(qrs_peer_rows_stage: steps per reference, scan, copy; ms = the whole stage, bytes = what came over NVLink)"
                                 if staged else
                                 "neighbour rows of other ranks, read in place by the prediction kernel (no separate exchange stage)")
            if not staged:
                rep[name]["ms"], rep[name]["nvlink_frac_of_770GBps"] = None, None
    rep["total_ms"] = sum(v for kx, v in stage_ms.items() if kx.startswith("exchange:"))
    rep["total_note"] = "time on the main stream: record all-to-all, pair all-to-all, exposed wait for the signature all-gather, barrier, staging of remote neighbour rows"
    return rep


def measure_extras(args, pipe, rowptr, colidx, val, n, m, nnz, peak, dev):
    """f-2 / f-3 / f-4 measured outside the timed step, on one extra step with int32 prediction columns."""
    import torch
    from query_recommendation_system_b200 import stages as _st
    saved = pipe.cols_mode
    pipe.cols_mode = 0
    try:
        out = pipe.run(rowptr, colidx, val)
        out = pipe.run(rowptr, colidx, val)
    finally:
        pipe.cols_mode = saved
    res = {}
    # ---------------------------------------------------------------- f-2: per-user top-10 recommendations
    if out.pred_row_start is not None:
        for _ in range(2):
            pipe.recommend_top_n(out, 10)
        t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0e.record()
        for _ in range(5):
            pipe.recommend_top_n(out, 10)
        t1e.record()
        torch.cuda.synchronize()
        res["recommend_top_n"] = {"ms": t0e.elapsed_time(t1e) / 5, "N": 10,
                                  "note": "qrs_recommend_topn over the sparse predictions of one step (outside the timed step)"}
    # ---------------------------------------------------------------- f-3: centred-cosine scoring of the step's pairs
    if out.pairs is not None and out.pairs.numel() > 0:
        pr = out.pairs.clone()
        P = pr.numel()
        bits = 32 + max(1, (n - 1).bit_length())

        def score_by_ratings():
            cval, n2c = _st.center_rows(rowptr, val)
            _st.sort_u64(pr, P, 1, 32, bits)
            return _st.pair_centered_cosine(rowptr, colidx, val, cval, n2c, pr, m, grouped=True)
        for _ in range(2):
            score_by_ratings()
        t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0e.record()
        for _ in range(5):
            score_by_ratings()
        t1e.record()
        torch.cuda.synchronize()
        ms_us = t0e.elapsed_time(t1e) / 5
        lens = (rowptr[1:] - rowptr[:-1])
        second = (pr & 0xFFFFFFFF)
        usim_bytes = float(lens[second].sum().item()) * 6 + nnz * 3 + P * (16 * ((bits - 32 + 7) // 8) + 12)
        res["user_similarity"] = {"ms": ms_us, "pairs": P, "pairs_per_s": P / (ms_us * 1e-3),
                                  "algorithmic_GBps": usim_bytes / (ms_us * 1e-3) / 1e9,
                                  "hbm_frac": usim_bytes / (ms_us * 1e-3) / 1e9 / peak,
                                  "note": "qrs_center_rows + qrs_sort_u64 + qrs_pair_centered_cosine on the candidate pairs of one "
                                          "step (user-branch scoring, recommender.py:263-269; outside the timed step)"}
    # ---------------------------------------------------------------- f-4: the CSV wire format either side of the path
    if m <= 196608 and n * (2 * m + 16) + 4 * nnz < 8e9:
        uids = ["U%d" % (i + 1) for i in range(n)]
        qids = ["Q%d" % (j + 1) for j in range(m)]

        def timed(fn, reps=3):
            fn()
            a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                r = fn()
            b_.record()
            torch.cuda.synchronize()
            return r, a.elapsed_time(b_) / reps
        enc = [u.encode() for u in uids]
        ids_off = torch.from_numpy(np.concatenate([[0], np.cumsum([len(e) for e in enc])]).astype(np.int64)).to(dev)
        ids = torch.from_numpy(np.frombuffer(b"".join(enc), dtype=np.uint8).copy()).to(dev)
        body, ms_fmt = timed(lambda: _st.csv_format(ids, ids_off, rowptr, colidx, val, m, None, True))
        header = torch.from_numpy(np.frombuffer((",".join(qids) + "\r\n").encode(), dtype=np.uint8).copy()).to(dev)
        text = torch.cat([header, body])

        def parse():
            ends = _st.csv_line_ends(text, max_lines=text.numel() // (m + 1) + 2)
            return _st.csv_ratings(text, ends, 1, ends.numel() - 1, m)
        (rp2, ci2, v2, _), ms_parse = timed(parse)
        same = bool(torch.equal(rp2, rowptr) and torch.equal(ci2, colidx) and torch.equal(v2, val))
        pred_text_ms = pred_bytes_txt = None
        if out.pred_row_start is not None:
            pred = (out.pred_row_start, out.pred_row_cnt, out.pred_cols, out.pred_vals)
            ptxt, pred_text_ms = timed(lambda: _st.csv_format(ids, ids_off, rowptr, colidx, val, m, pred, False), reps=2)
            pred_bytes_txt = int(ptxt.numel())
            del ptxt
        tb = float(text.numel())
        res["csv_io"] = {"utility_matrix_bytes": int(tb), "round_trip_identical": same,
                         "parse_ms": ms_parse, "parse_GBps": (tb + 5 * nnz + 8 * n) / (ms_parse * 1e-3) / 1e9,
                         "format_ms": ms_fmt, "format_GBps": (tb + 5 * nnz + 8 * n) / (ms_fmt * 1e-3) / 1e9,
                         "final_predictions_ms": pred_text_ms, "final_predictions_bytes": pred_bytes_txt,
                         "note": "utility_matrix.csv text formatted on the device, parsed back and compared; final_predictions.csv "
                                 "text = ratings + one step's predictions; outside the timed step"}
        del text, body
    return res


if __name__ == "__main__":
    main()
