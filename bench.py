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

`--impl reference` times the reference's CPU algorithm (oracle/port.py — /root/reference does not
exist on the GPU box and is pure Python, so kind = "port") on a bounded sample of the same workload.
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


def run_cpu_port(config, n_sets):
    from oracle import baseline
    (rowptr, colidx, val), (m, K, b, k) = cpu_sample(config, n_sets)
    t, counts = baseline.run_port_pipeline(rowptr, colidx, val, m, K, b, k, perm_seed=PERM_SEED)
    return t, counts


def reference_arm(args):
    """bench.py --impl reference: the reference's CPU path (port) on a bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = args.steps, args.warmup
    n_sets = int(max(8, min(64, (64 * 150.0 / (10.6 * (steps + warmup))) // 8 * 8)))
    for _ in range(warmup):
        run_cpu_port(args.config, n_sets)
    t0 = time.perf_counter()
    stage = {}
    for _ in range(steps):
        t, counts = run_cpu_port(args.config, n_sets)
        for kx, v in t.items():
            stage[kx] = stage.get(kx, 0.0) + v
    wall = time.perf_counter() - t0
    path_s = stage["total"] / steps
    value = n_sets / path_s
    sample = ("first %d sets of the %s matrix (closed sample: LSH and CF run over these sets only), whole path per "
              "step; 1 Python thread, like the reference" % (n_sets, args.config))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": path_s * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int64/f64", "data": "synthetic",
        "config": workload_config(args.config, 1, "n/a (CPU)"),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample,
                         "host_cores": os.cpu_count(),
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


def workload_config(config, n_gpus, l2_note):
    from query_recommendation_system_b200 import synth
    n, m, density, K, b, k = synth.config_shape(config)
    return {"workload": "%s: %d users x %d queries, %.3g%% density, K=%d hashes, b=%d bands x r=%d rows, top-k=%d, "
                        "planted taste groups of 32 (SURVEY.md 8d), ratings 1..100"
                        % (config, n, m, density * 100, K, b, K // b, k),
            "sets_per_gpu": n, "elements": m, "density": density, "K": K, "b": b, "r": K // b, "k": k,
            "parallelism": "row shards x%d, bucket all-to-all" % n_gpus if n_gpus > 1 else "single GPU",
            "binding": "B (set = user row of the utility matrix, element = rated query)",
            "permutations": "np.random.seed(%d); K x np.random.permutation(m) (setup, not timed)" % PERM_SEED,
            "l2": l2_note}


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


# ------------------------------------------------------------------------------------------ GPU side
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default="C2")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sets", type=int, default=160)
    ap.add_argument("--sweep", action="store_true", help="band/row sweep of the LSH stages (extra, not the bench line)")
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

    n, m, density, K, b, k = synth.config_shape(args.config)
    lib = _lib.load()
    rowptr, colidx, val = synth.planted_csr(n, m, density, seed=DATA_SEED, row_base=rank * n, device=dev)
    nnz = int(colidx.numel())
    tab = perm_table(m, K)
    pipe = LshCfPipeline(m=m, K=K, b=b, k=k, perm_tab=tab, device=dev, profile=True)
    input_bytes = rowptr.numel() * 8 + colidx.numel() * 4 + val.numel()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    warm = max(args.warmup, 3)
    for _ in range(warm):
        out = pipe.run(rowptr, colidx, val)
    barrier()

    launches0 = lib.qrs_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stage_ms = {}
    barrier()
    sampler.mark_start()
    ev0.record()
    outs = []
    for _ in range(args.steps):
        out = pipe.run(rowptr, colidx, val)
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

    # per-rank counts → whole-job totals
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
                    "note": "stage = tile-index kernel + predict kernel + one cursor readback"},
    }
    for v in rl.values():
        v["frac"] = v["achieved"] / peak
    dominant = max(("signatures+banding", "predict"), key=lambda s: stage_ms.get(s, 0.0))
    roofline = dict(rl[dominant])
    kernel_name = "predict_sparse_kernel" if dominant == "predict" else ("minhash_tile_kernel" if m <= 32768 else "minhash_kernel")
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r1_traffic.json")     # DRAM bytes per launch from the committed ncu capture
    if world == 1 and args.config == "C2" and os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(kernel_name)
    roofline.update({"kernel": kernel_name, "stage": dominant, "peak_source": peak_src + " (of measured)",
                     "traffic": traffic, "traffic_source": "profiles/r1_traffic.json (ncu --set full, same command)"})
    try:    # the on-chip roof the dominant kernel is actually close to (counter from the same committed capture)
        wf = json.load(open(tpath)).get(kernel_name + ":smem_wavefronts") if traffic is not None else None
        if wf:
            roofline["onchip"] = onchip_roofline(float(wf), roofline["ms"], torch.cuda.get_device_properties(dev).multi_processor_count,
                                                 float(clocks.get("sm_mhz") or 1965.0))
    except Exception:
        pass

    # ---------------------------------------------------------------- f-2 extra: per-user top-10 recommendations
    topn = None
    if world == 1 and out.pred_row_start is not None:
        for _ in range(2):
            pipe.recommend_top_n(out, 10)
        t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0e.record()
        for _ in range(5):
            pipe.recommend_top_n(out, 10)
        t1e.record()
        torch.cuda.synchronize()
        topn = {"ms": t0e.elapsed_time(t1e) / 5, "N": 10,
                "note": "qrs_recommend_topn over the sparse predictions of the last step (outside the timed step)"}

    # ---------------------------------------------------------------- f-3 extra: centred-cosine scoring of the step's pairs
    usim = None
    if world == 1 and out.pairs is not None and out.pairs.numel() > 0:
        from query_recommendation_system_b200 import stages as _st
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
        # algorithmic bytes: the second row of every pair is streamed once (4 B column + 2 B centred value per
        # cell), centring reads 1 B and writes 2 B per cell, the pair list is sorted (2 passes x 16 B) and read
        usim_bytes = float(lens[second].sum().item()) * 6 + nnz * 3 + P * (16 * ((bits - 32 + 7) // 8) + 12)
        usim = {"ms": ms_us, "pairs": P, "pairs_per_s": P / (ms_us * 1e-3),
                "algorithmic_GBps": usim_bytes / (ms_us * 1e-3) / 1e9, "hbm_frac": usim_bytes / (ms_us * 1e-3) / 1e9 / peak,
                "note": "qrs_center_rows + qrs_sort_u64 + qrs_pair_centered_cosine on the candidate pairs of the last "
                        "step (user-branch scoring, recommender.py:263-269; outside the timed step)"}

    # ---------------------------------------------------------------- f-4 extra: the CSV wire format either side of the path
    csvio = None
    if world == 1 and m <= 196608 and n * (2 * m + 16) + 4 * nnz < 8e9:
        from query_recommendation_system_b200 import io_csv as _io, stages as _st
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
        pred_text_ms = None
        if out.pred_row_start is not None:
            pred = (out.pred_row_start, out.pred_row_cnt, out.pred_cols, out.pred_vals)
            ptxt, pred_text_ms = timed(lambda: _st.csv_format(ids, ids_off, rowptr, colidx, val, m, pred, False), reps=2)
            pred_bytes_txt = int(ptxt.numel())
            del ptxt
        tb = float(text.numel())
        csvio = {"utility_matrix_bytes": int(tb), "round_trip_identical": same,
                 "parse_ms": ms_parse, "parse_GBps": (tb + 5 * nnz + 8 * n) / (ms_parse * 1e-3) / 1e9,
                 "format_ms": ms_fmt, "format_GBps": (tb + 5 * nnz + 8 * n) / (ms_fmt * 1e-3) / 1e9,
                 "final_predictions_ms": pred_text_ms,
                 "final_predictions_bytes": pred_bytes_txt if pred_text_ms is not None else None,
                 "hbm_peak_GBps": peak,
                 "note": "utility_matrix.csv text of this step's ratings formatted on the device "
                         "(qrs_csv_format_*), parsed back (qrs_csv_line_ends + qrs_csv_ratings_*) and compared; "
                         "final_predictions.csv text = ratings + the step's predictions; GB/s = (text + CSR bytes) / time; "
                         "outside the timed step"}
        del text, body

    runner = HostRunner(pipe)
    h_rowptr, h_colidx, h_val = [x.cpu().pin_memory() for x in (rowptr, colidx, val)]
    for _ in range(2):
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

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32/int64 (signatures, buckets, cosine dot) + f64 (similarity, prediction)",
            "data": "synthetic",
            "config": workload_config(args.config, world, "inputs (%.0f MB CSR per GPU) exceed the 126 MB L2; no flush"
                                      % (input_bytes / 1e6)),
            "stages": {
                "ms": stage_ms,
                "sigs_per_s": n_total / (ms_sig * 1e-3) if world == 1 else n / (ms_sig * 1e-3) * world,
                "candidate_pairs_per_s": n_pairs / (sum(stage_ms.get(s, 0.0) for s in ("grouping", "pairs+cosine")) * 1e-3),
                "pair_slots_per_s": n_slots / (sum(stage_ms.get(s, 0.0) for s in ("grouping", "pairs+cosine")) * 1e-3),
                "predicted_ratings_per_s": covered / (ms_pred * 1e-3),
                "nonzero_predictions_per_s": n_pred / (ms_pred * 1e-3),
                "counts": {"sets": n_total, "nnz_per_gpu": nnz, "candidate_pairs": n_pairs, "pair_slots": n_slots,
                           "nonzero_predictions": n_pred, "covered_cells": covered},
            },
            "roofline": roofline, "rooflines": rl,
            "e2e": {"value": n_total / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms, "wall_ms_per_step": e2e_wall_ms,
                    "h2d_bytes_per_step": runner.h2d_bytes, "d2h_bytes_per_step": runner.d2h_bytes, "steps": e2e_steps,
                    "call": "pipeline.HostRunner(pipe, chunks=4, cols_format=auto)(rowptr, colidx, val) with pinned host tensors; returns pinned "
                            "sparse predictions + neighbour tables"},
            "gpu_launches": int(launches), "clocks": clocks, "recommend_top_n": topn, "user_similarity": usim, "csv_io": csvio,
        }
        if world == 1 and not args.no_cpu_baseline:
            tc, cc = run_cpu_port(args.config, args.cpu_sets)
            line["cpu_baseline"] = {
                "value": args.cpu_sets / tc["total"], "unit": UNIT, "cores": 1, "kind": "port",
                "host_cores": os.cpu_count(), "seconds": tc, "counts": cc,
                "sample": "first %d sets of the %s matrix as a closed sample, whole path once; single Python thread "
                          "(the reference is single-threaded)" % (args.cpu_sets, args.config)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
