#!/usr/bin/env python
"""bench.py — query spectra/s through the cascade open-search hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repository's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle/)

A step is one pass of the hot path (vectorise -> IVF top-k -> precursor window -> shifted dot)
over one batch of synthetic queries per rank: BASELINE.json configs[1] ("C2": 16,384 queries vs
a 3 M-vector library, charges 2-4, nprobe 1024, top-1024, 500 Da open window). Weak scaling: every
rank holds a replica of the library/index (mode A) and its own 16,384-query batch.
`value` is timed with the batch already resident in HBM, `e2e` through the C-ABI with host
buffers (H2D of the queries and D2H of the results inside the timed region).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (n_targets, decoy_fraction, queries per rank, nlist, nprobe, k, cpu sample queries)
    "c2": dict(n_targets=2_000_000, decoys=0.5, nq=16384, nlist=16384, nprobe=1024, k=1024, cpu_sample=6144,
               label="C2: 16,384 queries vs 3M-vector library (2M targets + 1M decoys), charges 2-4"),
    "c2s": dict(n_targets=2_000_000, decoys=0.5, nq=16384, nlist=4096, nprobe=1024, k=1024, cpu_sample=128,
                label="C2 (second point, nlist 4096)"),
    # C5 point: 10 M vectors (mode B at N > 1: the `sharded` object; the library still fits one HBM, so mode A runs too)
    "c5": dict(n_targets=6_666_667, decoys=0.5, nq=16384, nlist=16384, nprobe=1024, k=1024, cpu_sample=0,
               label="C5 point: 16,384 queries vs 10M-vector library (6.67M targets + 3.33M decoys), charges 2-4"),
    "c1": dict(n_targets=200_000, decoys=0.5, nq=16384, nlist=256, nprobe=128, k=1024, cpu_sample=512,
               label="C1: 16,384 queries vs 200k+100k-decoy library, charges 2-4"),
    # streamed workloads (run_stream): the whole query set goes through the engine in batches of `nq`, like the
    # reference's loop over config.batch_size (spectral_library.py:301-306)
    "c3": dict(n_targets=2_000_000, decoys=0.5, nq=16384, nlist=16384, nprobe=1024, k=1024, cpu_sample=0,
               stream_queries=1_048_576, stream_scaling="strong", levels=("open",),
               label="C3: 1,048,576 queries (ONE set, partitioned over the ranks) vs 3M-vector library, streamed in "
                     "batches of 16,384 per charge, results gathered to rank 0"),
    "c4": dict(n_targets=2_000_000, decoys=0.5, nq=16384, nlist=16384, nprobe=1024, k=1024, cpu_sample=0,
               stream_queries=1_048_576, stream_scaling="weak", levels=("std", "open"),
               label="C4 (bounded): 64 batches of 16,384 queries per rank of the 25M-query run vs 3M-vector library, "
                     "both cascade levels (20 ppm window-only search, then the 500 Da ANN open search) on every query"),
    "c3s": dict(n_targets=20_000, decoys=0.5, nq=1024, nlist=64, nprobe=16, k=128, cpu_sample=0,
                stream_queries=10_000, stream_scaling="strong", levels=("std", "open"), label="streamed smoke workload"),
    "tiny": dict(n_targets=20_000, decoys=0.5, nq=1024, nlist=64, nprobe=16, k=128, cpu_sample=256,
                 label="tiny smoke workload"),
}
OPEN_TOL, OPEN_MODE, FRAG_TOL = 500.0, "Da", 0.02
TRAIN_ITERS = int(os.environ.get("SOLO_TRAIN_ITERS", "2"))  # Lloyd iterations of the index build (not timed)


_emit = print


def log(*a):
    if int(os.environ.get("RANK", "0")) == 0:
        print("[bench]", *a, file=sys.stderr, flush=True)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc:
            self.proc.terminate()
            self.t.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except (ValueError, IndexError):
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_data(wl, rank):
    from ann_solo_b200 import synth
    t0 = time.time()
    lib = synth.make_library(wl["n_targets"], decoy_fraction=wl["decoys"], seed=1, decoy_seed=2)
    per_charge = synth.split_by_charge(lib)
    queries = synth.make_queries(lib, wl["nq"], seed=3 + rank)
    q_by_charge = {}
    for z in sorted(per_charge):
        sel = np.flatnonzero(queries["prec_z"] == z)
        if len(sel):
            q_by_charge[z] = synth.take_spectra(queries, sel)
    log(f"synthetic data: {len(lib['prec_mz'])} library spectra, {wl['nq']} queries/rank in {time.time() - t0:.1f}s")
    return lib, per_charge, q_by_charge


def pinned_copy(torch, a):
    t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    return t, t.numpy()


def run_solo(args, wl, rank, world, local_rank):
    import torch
    from ann_solo_b200.engine import SoloEngine

    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib, per_charge, q_by_charge = make_data(wl, rank)
    # one non-default stream carries torch's timing events, the NCCL hand-offs and every kernel of the engine
    torch.cuda.set_stream(torch.cuda.Stream())
    eng = SoloEngine(local_rank)
    eng.set_stream(torch.cuda.current_stream().cuda_stream)
    for kv in filter(None, os.environ.get("SOLO_OPT", "").split(",")):  # tuning switches, e.g. round0_scores=6144
        key, val = kv.split("=")
        eng.set_option(key, int(val))
    t0 = time.time()
    charges = sorted(q_by_charge)
    n_vec = {}
    for z in charges:
        store, _ = per_charge[z]
        eng.load_library(z, store)
        nlist = min(wl["nlist"], max(1, len(store["prec_mz"]) // 39))
        eng.ivf_train_library(z, nlist, iters=TRAIN_ITERS, seed=4)
        eng.ivf_add_library(z)
        n_vec[z] = (len(store["prec_mz"]), nlist)
    eng.synchronize()
    log(f"index build (k-means {TRAIN_ITERS} it. + add, all charges): {time.time() - t0:.1f}s  {n_vec}")
    if os.environ.get("SOLO_BENCH_STATS") == "1":
        for z in charges:
            a_ = eng.ivf_assignment(z)
            sizes = np.bincount(a_[a_ >= 0], minlength=n_vec[z][1])
            log(f"list sizes z={z}: mean {sizes.mean():.1f} p50 {np.percentile(sizes, 50):.0f} p90 {np.percentile(sizes, 90):.0f} "
                f"max {sizes.max()} empty {(sizes == 0).sum()} frac>96 {(sizes > 96).mean():.3f} frac>192 {(sizes > 192).mean():.3f} "
                f"chunks/list {np.ceil(sizes / 96).mean():.3f}")

    max_pairs = 50
    params = SoloEngine.make_params(True, wl["k"], wl["nprobe"], OPEN_TOL, OPEN_MODE, FRAG_TOL, True, max_pairs)
    # pinned host copies of the queries and of the result buffers (e2e path)
    pin_keep, host_q, host_out = [], {}, {}
    h2d_bytes = d2h_bytes = 0
    for z in charges:
        q = q_by_charge[z]
        hq = {}
        for key in ("mz", "inten", "off", "prec_mz"):
            t, hq[key] = pinned_copy(torch, q[key])
            pin_keep.append(t)
            h2d_bytes += hq[key].nbytes
        host_q[z] = hq
        nq = len(q["prec_mz"])
        out = {}
        for key, shape, dt in (("best_row", (nq,), np.int32), ("score", (nq,), np.float64),
                               ("n_pairs", (nq,), np.int32), ("pairs", (nq, max_pairs, 2), np.uint32),
                               ("n_cand", (nq,), np.int32)):
            t = torch.empty(shape, dtype=getattr(torch, np.dtype(dt).name.replace("uint32", "int32"))).pin_memory()
            pin_keep.append(t)
            out[key] = t.numpy().view(dt)
            d2h_bytes += out[key].nbytes
        host_out[z] = out
    nq_rank = sum(len(q_by_charge[z]["prec_mz"]) for z in charges)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: inputs resident in HBM (one slot per charge), compute only
    for z in charges:
        eng.select_slot(z)
        eng.stage_queries(host_q[z])
    eng.synchronize()

    def step_resident():
        for z in charges:
            eng.select_slot(z)
            eng.search_staged(z, params)

    def step_e2e(n_steps=1):
        # the public streaming call: pinned host buffers in, pinned host results out; the H2D of the next batch and
        # the D2H of the previous one run on the copy stream under the kernels (SoloEngine.search_stream). The K
        # timed steps are one stream of 3 K batches, like the reference's loop over config.batch_size batches.
        for _ in eng.search_stream(params, [(z, host_q[z], host_out[z]) for _ in range(n_steps) for z in charges]):
            pass

    def timed(step_fn, steps, warmup, profile=False, streamed=False):
        for _ in range(warmup):
            step_fn(2) if streamed else step_fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ncu = profile and os.environ.get("SOLO_NCU") == "1"  # ncu --profile-from-start off
        with ClockSampler(local_rank) as clk:
            # nvidia-smi polls every 200 ms and the K timed steps may be shorter than that: keep the GPU
            # under the same load for about a second first so that samples are taken under load
            t_load = time.time()
            while not ncu and time.time() - t_load < 1.0:
                step_fn()
                torch.cuda.synchronize()
            barrier()
            if profile:
                eng.profile_reset()
                eng.profile_enable(True)
            launches0 = eng.kernel_launches()
            if ncu:
                torch.cuda.profiler.start()
            e0.record()
            if streamed:
                step_fn(steps)
            else:
                for _ in range(steps):
                    step_fn()
            e1.record()
            barrier()
            if ncu:
                torch.cuda.profiler.stop()
        ms = e0.elapsed_time(e1)
        prof = eng.profile() if profile else None
        eng.profile_enable(False)
        launches = eng.kernel_launches() - launches0
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), prof, launches, clk.summary()

    if os.environ.get("SOLO_BENCH_STATS") == "1":  # diagnostic: per-query scan buffer fill and candidate counts
        import ctypes as C
        for z in charges:
            eng.select_slot(z)
            eng.search_staged(z, params)
            eng.synchronize()
            nq = len(q_by_charge[z]["prec_mz"])
            cap, cnt = C.c_int32(), np.empty(nq, np.int32)
            eng._check(eng._lib.solo_debug_scan_dump(eng._h, int(z), nq, C.byref(cap), cnt.ctypes.data_as(C.c_void_p), None))
            eng._staged_nq = nq
            nc = eng.fetch_results()["n_cand"]
            pc = np.percentile(cnt, [1, 50, 90, 99, 100]).astype(int).tolist()
            log(f"stats z={z}: nq={nq} scan-buffer entries/query p1/p50/p90/p99/max = {pc} mean={cnt.mean():.0f}; "
                f"scored candidates/query mean={nc.mean():.0f} max={nc.max()}")
    for variant in filter(None, os.environ.get("SOLO_BENCH_SWEEP", "").split(";")):  # diagnostic: option sweep
        for kv in variant.split(","):
            key, val = kv.split("=")
            eng.set_option(key, int(val))
        ms_v, prof_v, _, _ = timed(step_resident, args.steps, args.warmup, profile=True)
        log(f"sweep [{variant}] {ms_v / args.steps:.2f} ms/step " +
            " ".join(f"{k}={v['ms'] / args.steps:.2f}" for k, v in prof_v.items() if v["ms"] > 0))
        if os.environ.get("SOLO_BENCH_STATS") == "1":
            import ctypes as C
            for z in charges:
                eng.select_slot(z)
                eng.search_staged(z, params)
                eng.synchronize()
                nq = len(q_by_charge[z]["prec_mz"])
                cap, cnt = C.c_int32(), np.empty(nq, np.int32)
                eng._check(eng._lib.solo_debug_scan_dump(eng._h, int(z), nq, C.byref(cap), cnt.ctypes.data_as(C.c_void_p), None))
                pc = np.percentile(cnt, [1, 50, 90, 99, 100]).astype(int).tolist()
                log(f"   z={z}: scan-buffer entries/query p1/p50/p90/p99/max = {pc} mean={cnt.mean():.0f}")
    ms_res, prof, launches, clocks = timed(step_resident, args.steps, args.warmup, profile=True)
    e2e_prof = os.environ.get("SOLO_BENCH_E2E_PROFILE") == "1"   # diagnostic: stage times of the streamed leg
    ms_e2e, prof_e2e, _, _ = timed(step_e2e, args.steps, max(1, args.warmup // 2), profile=e2e_prof, streamed=True)
    if e2e_prof:
        log("e2e stages (ms/step): " + " ".join(f"{k}={v['ms'] / args.steps:.2f}" for k, v in prof_e2e.items() if v["ms"] > 0))
    if os.environ.get("SOLO_BENCH_E2E_SYNC") == "1":             # diagnostic: the synchronous one-call API instead
        def step_sync():
            for z in charges:
                eng.select_slot(z)
                eng.search_batch(z, params, host_q[z], out=host_out[z])
        ms_sync, _, _, _ = timed(step_sync, args.steps, 1)
        log(f"e2e through the synchronous solo_search_batch: {ms_sync / args.steps:.3f} ms/step (streamed: {ms_e2e / args.steps:.3f})")
    log("resident stages (ms/step): " + " ".join(f"{k}={v['ms'] / args.steps:.2f}" for k, v in prof.items() if v["ms"] > 0))
    value = world * nq_rank * args.steps / (ms_res / 1e3)
    e2e = world * nq_rank * args.steps / (ms_e2e / 1e3)

    # ---- sanity: the e2e results of the last step are real
    n_match = int(sum((host_out[z]["best_row"] >= 0).sum() for z in charges))
    assert n_match > 0.5 * nq_rank, f"only {n_match} of {nq_rank} queries matched"

    # ---- mode B on the same library (N > 1, or --sharded): every rank takes part, rank 0 reports
    sharded = None
    if world > 1 or args.sharded:
        sharded = measure_sharded(args, wl, rank, world, eng, charges, lib, params, torch, dist,
                                  check_against=host_out if rank == 0 else None)
    if rank != 0:
        return
    # ---- roofline of the dominant kernel (K3 list scan): algorithmic flops 2*d*S per query
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_src = "measured (MEASURED_PEAKS.json bf16_tflops_sustained)" if peaks else "fallback (B200_PROFILING.md)"
    scan = prof["scan"]
    scan_ms_per_launch = scan["ms"] / max(scan["launches"], 1)
    flops_per_launch = scan["units"] / max(scan["launches"], 1)
    achieved = flops_per_launch / (scan_ms_per_launch * 1e-3) / 1e12 if scan_ms_per_launch > 0 else 0.0
    # DRAM bytes (read + write) per launch of the dominant kernel from the committed `ncu --set full` capture
    traffic, tr_all = None, {}
    try:   # written by profiles/export_ncu.py from the committed `ncu --set full` capture (carries the build's git hash)
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(args.workload, {})
        traffic = tr.get("k3_list_scan_dram_bytes_per_launch")
        tr_all = tr.get("per_kernel_dram_bytes_per_launch", {})
    except (OSError, ValueError):
        pass
    roofline = {"kernel": "k3_list_scan", "bound": "tensor", "achieved": round(achieved, 3), "peak": peak_tf,
                "unit": "TFLOP/s", "frac": round(achieved / peak_tf, 5), "traffic": traffic, "peak_source": peak_src,
                "ms_per_launch": round(scan_ms_per_launch, 4), "launches": scan["launches"],
                "share_of_step": round(scan["ms"] / ms_res, 4)}
    stages = {k: round(v["ms"] / args.steps, 4) for k, v in prof.items() if v["ms"] > 0}
    # the other two kernels with a roofline of their own: K2 coarse scoring (tensor: 2 d nlist flops per
    # query) and K5 (HBM gather: 9 P_c + 24 bytes per scored (query, candidate) pair, P_c = peaks of the candidate)
    hbm_gbs = peaks.get("hbm_gbs_sustained", peaks.get("hbm_gbs", 6500.0))
    coarse = prof["coarse"]
    coarse_tf = coarse["units"] / (coarse["ms"] * 1e-3) / 1e12 if coarse["ms"] > 0 else 0.0
    pairs_scored = float(sum(int(host_out[z]["n_cand"].sum()) for z in charges))
    mean_pc = float(np.mean([np.diff(per_charge[z][0]["off"]).mean() for z in charges]))
    k5_bytes = pairs_scored * (9.0 * mean_pc + 24.0)
    k5_ms = prof["score"]["ms"] / args.steps
    k5_gbs = k5_bytes / (k5_ms * 1e-3) / 1e9 if k5_ms > 0 else 0.0
    roofline_all = [
        dict(roofline),
        {"kernel": "k2_coarse (thresholded tcgen05 pass incl. the sampled prefix)", "bound": "tensor", "achieved": round(coarse_tf, 3),
         "peak": peak_tf, "unit": "TFLOP/s", "frac": round(coarse_tf / peak_tf, 5), "ms_per_step": round(coarse["ms"] / args.steps, 4),
         "traffic": tr_all.get("k2_coarse")},
        {"kernel": "k5_fast_kernel (shifted dot)", "bound": "hbm", "achieved": round(k5_gbs, 1), "peak": hbm_gbs, "unit": "GB/s",
         "frac": round(k5_gbs / hbm_gbs, 5), "ms_per_step": round(k5_ms, 4), "pairs_per_step": int(pairs_scored),
         "algorithmic_bytes_per_pair": round(9.0 * mean_pc + 24.0, 1), "traffic": tr_all.get("k5_fast_kernel")},
    ]

    # the CPU baseline is reported by the single-GPU run only (torchrun also pins OMP_NUM_THREADS=1)
    cpu, parity = (cpu_baseline(wl, per_charge, q_by_charge, eng, host_out)
                   if (not args.no_cpu_baseline and world == 1) else (None, None))
    extras = measure_extras(eng, charges, q_by_charge, nq_rank, torch) if args.extras else None
    line = {
        "metric": "query spectra/sec, cascade open search", "value": round(value, 1), "unit": "spectra/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_res / args.steps, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32/f64", "data": "synthetic",
        "config": {"workload": wl["label"], "queries_per_rank_per_step": nq_rank, "nlist": wl["nlist"],
                   "nprobe": wl["nprobe"], "k": wl["k"], "open_window": f"{OPEN_TOL} {OPEN_MODE}",
                   "fragment_tol": FRAG_TOL, "parallelism": f"queries partitioned x{world}, library replicated",
                   "l2": "index and peak store (>1 GB) exceed the 126 MB L2; no explicit flush"},
        "e2e": {"value": round(e2e, 1), "unit": "spectra/s", "h2d_bytes_per_step": int(h2d_bytes),
                "d2h_bytes_per_step": int(d2h_bytes), "ms_per_step": round(ms_e2e / args.steps, 3)},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "roofline_all": roofline_all,
        "stage_ms_per_step": stages, "cpu_baseline": cpu, "parity": parity,
    }
    if extras:
        line["extras"] = extras
    if sharded:
        line["sharded"] = sharded
    _emit(json.dumps(line))
    if parity and parity["mismatch"]:
        log(f"PARITY FAILURE: {parity['mismatch']} of {parity['checked']} sampled queries differ from the CPU reference path")
        sys.exit(1)


def run_stream(args, wl, rank, world, local_rank):
    """C3 / C4: a long query set streamed through SoloEngine.search_stream (copy stream + two query slots). One step =
    one pass over the whole set: every rank streams its share batch by batch (per charge, `nq` queries per batch)
    through every cascade level asked for, from pinned host buffers into pinned host results; with N > 1 the results
    are then gathered to rank 0 over NCCL inside the timed region."""
    import torch
    from ann_solo_b200 import parallel, synth
    from ann_solo_b200.engine import SoloEngine
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_stream(torch.cuda.Stream())
    t0 = time.time()
    lib = synth.make_library(wl["n_targets"], decoy_fraction=wl["decoys"], seed=1, decoy_seed=2)
    per_charge = synth.split_by_charge(lib)
    strong = wl["stream_scaling"] == "strong"
    queries = synth.make_queries(lib, wl["stream_queries"], seed=3 if strong else 3 + rank)
    log(f"synthetic data: {len(lib['prec_mz'])} library spectra, {wl['stream_queries']} queries in {time.time() - t0:.1f}s")
    eng = SoloEngine(local_rank)
    eng.set_stream(torch.cuda.current_stream().cuda_stream)
    charges = sorted(per_charge)
    t0 = time.time()
    for z in charges:
        store, _ = per_charge[z]
        eng.load_library(z, store)
        eng.ivf_train_library(z, min(wl["nlist"], max(1, len(store["prec_mz"]) // 39)), iters=TRAIN_ITERS, seed=4)
        eng.ivf_add_library(z)
    eng.synchronize()
    log(f"index build: {time.time() - t0:.1f}s")
    del lib
    max_pairs = 50
    level_params = {"open": SoloEngine.make_params(True, wl["k"], wl["nprobe"], OPEN_TOL, OPEN_MODE, FRAG_TOL, True, max_pairs),
                    "std": SoloEngine.make_params(False, wl["k"], wl["nprobe"], 20.0, "ppm", FRAG_TOL, True, max_pairs)}
    # this rank's share, cut into batches; pinned inputs and outputs
    pin_keep, batches, h2d_bytes, d2h_bytes, n_local = [], [], 0, 0, 0
    out_fields = (("best_row", (), np.int32), ("score", (), np.float64), ("n_pairs", (), np.int32),
                  ("pairs", (max_pairs, 2), np.uint32), ("n_cand", (), np.int32))
    for z in charges:
        rows = np.flatnonzero(queries["prec_z"] == z)
        if strong:
            b, e = parallel.shard_bounds(len(rows), rank, world)
            rows = rows[b:e]
        for b0 in range(0, len(rows), wl["nq"]):
            q = synth.take_spectra(queries, rows[b0:b0 + wl["nq"]])
            hq = {}
            for key in ("mz", "inten", "off", "prec_mz"):
                t, hq[key] = pinned_copy(torch, q[key])
                pin_keep.append(t)
                h2d_bytes += hq[key].nbytes
            n = len(q["prec_mz"])
            out = {}
            for key, tail, dt in out_fields:
                t = torch.empty((n,) + tail, dtype=getattr(torch, np.dtype(dt).name.replace("uint32", "int32"))).pin_memory()
                pin_keep.append(t)
                out[key] = t.numpy().view(dt)
                d2h_bytes += out[key].nbytes
            batches.append((z, hq, out))
            n_local += n
    del queries
    levels = wl["levels"]
    n_max = n_local
    if dist is not None:
        t = torch.tensor([n_local], dtype=torch.int64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        n_max = int(t.item())
    gathered = {}

    pinned_gather = {}

    def gather_to_rank0():
        # results of the (last) level to rank 0: one padded device tensor per field, gathered over NCCL, landed in
        # pinned host memory on rank 0
        for key, tail, dt in out_fields:
            tdt = getattr(torch, np.dtype(dt).name.replace("uint32", "int32"))
            # (padding rows behind a rank's own count read "no match")
            loc = torch.full((n_max,) + tail, -1 if key == "best_row" else 0, dtype=tdt, device="cuda")
            off = 0
            for _, _, o in batches:   # pinned result buffers -> device, asynchronously
                t = torch.from_numpy(o[key].view(np.dtype(dt).name.replace("uint32", "int32")))
                loc[off:off + len(t)].copy_(t, non_blocking=True)
                off += len(t)
            parts = [torch.empty_like(loc) for _ in range(world)] if rank == 0 else None
            dist.gather(loc, parts, dst=0)
            if rank == 0:
                if key not in pinned_gather:
                    pinned_gather[key] = torch.empty((world, n_max) + tail, dtype=tdt).pin_memory()
                for r in range(world):
                    pinned_gather[key][r].copy_(parts[r], non_blocking=True)
        torch.cuda.synchronize()
        if rank == 0:
            gathered.update(pinned_gather)

    def one_pass(sel=None):
        todo = batches if sel is None else batches[:sel]
        for lv in levels:
            for _ in eng.search_stream(level_params[lv], todo):
                pass
        if dist is None or sel is not None:
            return
        gather_to_rank0()

    one_pass(sel=min(6, len(batches)))     # warm-up on a few batches of the stream (buffers, slots, lazy module loads)
    for _ in range(max(0, args.warmup - 1)):
        one_pass(sel=min(6, len(batches)))
    if dist is not None:
        gather_to_rank0()   # warm-up of the collective too (NCCL sets its point-to-point channels up on first use)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        eng.profile_reset()
        eng.profile_enable(True)
        l0 = eng.kernel_launches()
        e0.record()
        for _ in range(args.steps):
            one_pass()
        e1.record()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
    prof = eng.profile()
    eng.profile_enable(False)
    launches = eng.kernel_launches() - l0
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    tot = torch.tensor([n_local, sum(int((o["best_row"] >= 0).sum()) for _, _, o in batches)], dtype=torch.int64, device="cuda")
    if dist is not None:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot)
    if rank != 0:
        return
    ms = float(ms.item())
    n_all, n_match = int(tot[0].item()), int(tot[1].item())
    assert n_match > 0.5 * n_all, f"only {n_match} of {n_all} queries matched"
    if gathered:
        assert int((gathered["best_row"] >= 0).sum()) == n_match, "gathered results differ from the ranks' own counts"
    value = n_all * args.steps / (ms / 1e3)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    scan = prof["scan"]
    achieved = scan["units"] / (scan["ms"] * 1e-3) / 1e12 if scan["ms"] > 0 else 0.0
    _emit(json.dumps({
        "metric": "query spectra/sec, cascade open search", "value": round(value, 1), "unit": "spectra/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True,
        "scaling": wl["stream_scaling"], "vs_baseline": None, "dtype": "f32/f64", "data": "synthetic",
        "config": {"workload": wl["label"], "queries_per_step": n_all, "batch": wl["nq"], "batches_rank0": len(batches),
                   "cascade_levels": list(levels), "nlist": wl["nlist"], "nprobe": wl["nprobe"], "k": wl["k"],
                   "parallelism": f"queries partitioned x{world}, library replicated" +
                                  (", results gathered to rank 0 (NCCL) inside the timed region" if world > 1 else ""),
                   "l2": "index and peak store (>1 GB) exceed the 126 MB L2; no explicit flush"},
        "e2e": {"value": round(value, 1), "unit": "spectra/s", "h2d_bytes_per_step": int(h2d_bytes * len(levels)),
                "d2h_bytes_per_step": int(d2h_bytes * len(levels)),
                "note": "the streamed run IS the end-to-end path: pinned host buffers in, pinned host results out, copies on "
                        "the copy stream under the kernels"},
        "gpu_launches": int(launches), "clocks": clk.summary(),
        "roofline": {"kernel": "k3_list_scan", "bound": "tensor", "achieved": round(achieved, 3), "peak": peak_tf,
                     "unit": "TFLOP/s", "frac": round(achieved / peak_tf, 5), "traffic": None,
                     "share_of_step": round(scan["ms"] / ms, 4)},
        "stage_ms_per_step": {k: round(v["ms"] / args.steps, 3) for k, v in prof.items() if v["ms"] > 0},
        "matched_queries": n_match, "cpu_baseline": None,
    }))


def measure_extras(eng, charges, q_by_charge, nq_rank, torch):
    """--extras: the widened rows of SURVEY.md §8f next to the hot path, on the same resident data.
    N4: K6 SSM feature table for the SSMs the last step produced (every charge), through the C-ABI with
    the (n, 44) float64 table copied to the host. N2: faiss.write_index / read_index of the smallest
    charge's index (float32 codes) to and from local disk."""
    import tempfile
    out = {}
    n_ssm, reps = 0, 5
    for z in charges:  # warm-up + count
        eng.select_slot(z)
        t = eng.ssm_features_staged(z)
        n_ssm += int(np.isfinite(t[:, 11]).sum())
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        for z in charges:
            eng.select_slot(z)
            eng.ssm_features_staged(z)
    dt = (time.perf_counter() - t0) / reps
    out["k6_ssm_features"] = {"ssm_per_s": round(n_ssm / dt, 1), "ms_per_batch": round(dt * 1e3, 3), "ssms": n_ssm,
                              "columns": 44, "d2h_bytes_per_batch": int(nq_rank * 44 * 8),
                              "note": "staged batch of every charge; wall clock around the synchronous C-ABI calls"}
    # N3: K0 process_spectrum over a batch of raw spectra (16,384 x ~300 peaks, gamma intensities), host buffers
    rng = np.random.default_rng(7)
    counts = rng.integers(100, 500, 16384)
    roff = np.zeros(len(counts) + 1, np.int64)
    np.cumsum(counts, out=roff[1:])
    rmz = rng.uniform(50.0, 2000.0, roff[-1]).astype(np.float32)
    rmz = rmz[np.lexsort((rmz, np.repeat(np.arange(len(counts)), counts)))]   # ascending inside every spectrum
    raw = dict(mz=rmz, inten=rng.gamma(0.6, 1000.0, roff[-1]).astype(np.float32), off=roff,
               prec_mz=rng.uniform(300, 1000, len(counts)), prec_z=rng.integers(2, 5, len(counts)).astype(np.int32))
    eng.process_spectra(raw)
    t0 = time.perf_counter()
    for _ in range(reps):
        done = eng.process_spectra(raw)
    dt = (time.perf_counter() - t0) / reps
    out["k0_process_spectrum"] = {"spectra_per_s": round(len(counts) / dt, 1), "ms_per_batch": round(dt * 1e3, 3),
                                  "raw_peaks": int(roff[-1]), "valid": int(done["valid"].sum()),
                                  "note": "16,384 raw spectra, host buffers in and out, incl. the NumPy compaction"}
    z = min(charges, key=lambda c: eng.ivf_info(c)[0])
    ntotal, nlist, d = eng.ivf_info(z)
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, f"bench_{z}.idxann")
        t0 = time.perf_counter()
        eng.ivf_write_index(z, path)
        t_w = time.perf_counter() - t0
        size = os.path.getsize(path)
        t0 = time.perf_counter()
        eng.ivf_read_index(9000 + z, path)
        t_r = time.perf_counter() - t0
        same = bool(np.array_equal(eng.ivf_assignment(9000 + z), eng.ivf_assignment(z)))
        eng.ivf_reset(9000 + z)
    out["idxann"] = {"charge": int(z), "vectors": int(ntotal), "nlist": int(nlist), "file_bytes": int(size),
                     "write_s": round(t_w, 3), "write_gbs": round(size / t_w / 1e9, 3), "read_s": round(t_r, 3),
                     "read_gbs": round(size / t_r / 1e9, 3), "assignment_identical_after_read": same}
    return out


def measure_sharded(args, wl, rank, world, eng, charges, lib, params, torch, dist, check_against=None):
    """Mode B on the same library: the inverted lists of every charge are dealt to the ranks by
    stored-vector count, every rank holds ONE global query batch (rank 0's), coarse scoring is sharded by queries
    (all-gather of the probe rows), the list scan by lists, the per-rank top-k rows reach the owner of each query
    slice with one all-to-all, are merged on the device and finished there (parallel.search_batch_sharded).
    Strong scaling: the global batch is fixed. Host query buffers are staged and the slice's results fetched every
    step. Returns the "sharded" object on rank 0 (None elsewhere)."""
    from ann_solo_b200 import parallel, synth
    queries = synth.make_queries(lib, wl["nq"], seed=3)           # rank 0's batch of the mode-A run, on every rank
    q_by_charge = {z: synth.take_spectra(queries, np.flatnonzero(queries["prec_z"] == z)) for z in charges}
    for z in charges:
        assign = eng.ivf_assignment(z)
        nl = eng.ivf_info(z)[1]
        sizes = np.bincount(assign[assign >= 0], minlength=nl)
        owner = parallel.assign_lists(sizes, world)
        eng.ivf_set_owned_lists(z, (owner == rank).astype(np.uint8))
    nq_total = sum(len(q_by_charge[z]["prec_mz"]) for z in charges)
    stats, last = {}, {}

    def step():
        sent = 0
        for z in charges:
            res = parallel.search_batch_sharded(eng, z, params, q_by_charge[z], rank, world, stats=stats)
            last[z] = res
            sent += stats.get("bytes_sent_per_rank", 0)
        return sent

    for _ in range(max(args.warmup, 2)):
        step()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    eng.profile_reset()
    eng.profile_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sent = 0
    for _ in range(args.steps):
        sent = step()
    e1.record()
    torch.cuda.synchronize()
    prof = eng.profile()
    eng.profile_enable(False)
    matched = sum(int((last[z]["best_row"] >= 0).sum()) for z in charges)
    mismatch = None
    if check_against is not None:      # rank 0: its slice against the replicated (mode A) results of the same batch
        mismatch = 0
        for z in charges:
            b, e, _ = parallel.slice_bounds(len(q_by_charge[z]["prec_mz"]), rank, world)
            for key in ("best_row", "n_cand", "n_pairs"):
                mismatch += int((last[z][key] != check_against[z][key][b:e]).sum())
            mismatch += int((last[z]["score"].view(np.uint64) != check_against[z]["score"][b:e].view(np.uint64)).sum())
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    m = torch.tensor([matched], dtype=torch.int64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(m)
    for z in charges:
        eng.ivf_set_owned_lists(z, None)
    if rank != 0:
        return None
    ms = float(t.item())
    return {"value": round(nq_total * args.steps / (ms / 1e3), 1), "unit": "spectra/s", "ms_per_step": round(ms / args.steps, 3),
            "scaling": "strong", "global_batch": nq_total, "matched_queries": int(m.item()),
            "bytes_sent_per_rank_per_step": int(sent),
            "exchange": "all-gather of probe rows (int32) + all-to-all of packed top-k entries (f32 score | u32 row) over NCCL",
            "stage_ms_per_step_rank0": {k: round(v["ms"] / args.steps, 4) for k, v in prof.items() if v["ms"] > 0},
            "mismatch_vs_replicated": mismatch,
            "note": "inverted lists sharded x%d; host query buffers staged and the slice's results fetched every step" % world}


def cpu_pipeline(o, store, q, cent, assign, nlist, wl, charge, threads, use_ref, simd=True):
    """The reference's CPU path for one batch of one charge: vectorise, IVF-Flat search (dense
    SIMD scan like Faiss), post-top-k window mask, SpectrumMatcher::dot (the reference's own
    compiled core when oracle/_ref exists)."""
    qv = o.vectorize(q["mz"], q["inten"], q["off"])
    _, ann = o.ivf_search(qv, cent, store["_list_off"], store["_list_ids"], store["_list_vecs"],
                          min(wl["nprobe"], nlist), wl["k"], simd=simd, n_threads=threads)
    cand, coff = o.candidates(q["prec_mz"], store["prec_mz"].astype(np.float32), store["valid"], charge, OPEN_TOL,
                              OPEN_MODE, ann)
    if use_ref:
        bp, bs, npairs, pairs = o.ref_best_match_batch(q, store, cand, coff, FRAG_TOL, True, n_threads=threads)
    else:
        bp, bs, npairs, pairs = o.best_match_batch(q, store, cand, coff, FRAG_TOL, True, sort_mode=0, n_threads=threads)
    row = np.full(len(bp), -1, np.int32)
    has = bp >= 0
    row[has] = cand[coff[:-1][has] + bp[has]]
    return {"best_row": row, "score": bs, "n_pairs": npairs, "pairs": pairs, "n_cand": np.diff(coff).astype(np.int32)}


def parity_mismatches(cpu, gpu, n):
    """Queries (of the first n) whose GPU result differs from the CPU reference path's: library row of the
    best match, score bits, candidate count, number of matched peak pairs and the pair set itself
    (canonically sorted: std::sort leaves the order inside exact product ties unspecified)."""
    bad = []
    for i in range(n):
        ok = (cpu["best_row"][i] == gpu["best_row"][i] and cpu["n_cand"][i] == gpu["n_cand"][i])
        if ok and cpu["best_row"][i] >= 0:
            m = int(cpu["n_pairs"][i])
            ok = (m == int(gpu["n_pairs"][i]) and
                  np.float64(cpu["score"][i]).view(np.uint64) == np.float64(gpu["score"][i]).view(np.uint64))
            if ok and m:
                a = np.asarray(cpu["pairs"][i, :m], np.int64)
                b = np.asarray(gpu["pairs"][i, :m], np.int64)
                ok = np.array_equal(a[np.lexsort((a[:, 1], a[:, 0]))], b[np.lexsort((b[:, 1], b[:, 0]))])
        if not ok:
            bad.append(i)
    return bad


def cpu_baseline(wl, per_charge, q_by_charge, eng, gpu_out=None):
    """Bounded sample of the same workload on the host cores, same centroids/lists as the GPU. When the
    GPU results of the same batch are given (`gpu_out[z]`, the e2e leg's host buffers), the sampled queries
    are compared with them: returns (cpu_baseline object, parity object)."""
    from ann_solo_b200 import synth
    from oracle import solo_oracle as o  # cpu_baseline leg only
    threads = o.num_threads()
    use_ref = o.have_ref()
    total_q, total_s = 0, 0.0
    checked = mismatch = resolved = 0
    n_all = sum(len(q["prec_mz"]) for q in q_by_charge.values())
    for z, q in q_by_charge.items():
        take = max(8, int(round(wl["cpu_sample"] * len(q["prec_mz"]) / n_all)))
        qs = synth.take_spectra(q, np.arange(min(take, len(q["prec_mz"]))))
        store = dict(per_charge[z][0])
        cent = eng.ivf_get_centroids(z)
        assign = eng.ivf_assignment(z)
        x = o.vectorize(store["mz"], store["inten"], store["off"])
        store["_list_off"], store["_list_ids"], store["_list_vecs"] = o.build_lists(x, assign, len(cent))
        del x
        t0 = time.perf_counter()
        res = cpu_pipeline(o, store, qs, cent, assign, len(cent), wl, z, threads, use_ref)
        total_s += time.perf_counter() - t0
        total_q += len(qs["prec_mz"])
        if gpu_out is not None:
            bad = parity_mismatches(res, gpu_out[z], len(qs["prec_mz"]))
            n_simd = len(bad)
            if bad:
                # the timed CPU scan sums in SIMD partial sums (like Faiss' fvec_inner_product), which can flip
                # the order of two near-equal scores at the k-th boundary; the parity definition is the
                # sequential-fmaf oracle, so these queries are looked at again with it
                sub = synth.take_spectra(qs, np.asarray(bad))
                again = cpu_pipeline(o, store, sub, cent, assign, len(cent), wl, z, threads, use_ref, simd=False)
                gsub = {k_: gpu_out[z][k_][np.asarray(bad)] for k_ in ("best_row", "score", "n_pairs", "pairs", "n_cand")}
                bad = parity_mismatches(again, gsub, len(bad))
            checked += len(qs["prec_mz"])
            mismatch += len(bad)
            resolved += n_simd - len(bad)
            log(f"parity z={z}: {len(qs['prec_mz'])} sampled queries vs the GPU results of the same batch: "
                f"{len(bad)} mismatches ({n_simd} before the sequential-fmaf re-check)")
        del store
    parity = None if gpu_out is None else {
        "checked": int(checked), "mismatch": int(mismatch), "simd_order_rechecked": int(resolved),
        "fields": "best library row, score bits (f64), candidate count, matched-pair count, canonical pair set",
        "against": "reference SpectrumMatch.cpp (oracle/_ref)" if use_ref else "ported scorer (oracle)"}
    return {"value": round(total_q / total_s, 2), "unit": "spectra/s", "cores": threads,
            "kind": "reference" if use_ref else "port",
            "sample": f"{total_q} queries of the same workload (all charges), vectorise + restated Faiss IVF-Flat "
                      f"(dense SIMD scan, OpenMP) + window + {'reference SpectrumMatch.cpp' if use_ref else 'ported scorer'}"
                      f" on {threads} threads, {total_s:.1f}s"}, parity


def run_reference(args, wl, rank, world):
    """--impl reference: the reference's CPU implementation of the path on the host cores (rank 0
    only). The index is built on the CPU too (same seeded k-means recipe, fast exact assignment)."""
    if rank != 0:
        return
    from ann_solo_b200 import synth
    from oracle import solo_oracle as o  # reference arm
    threads = o.num_threads()
    use_ref = o.have_ref()
    lib, per_charge, q_by_charge = make_data(wl, 0)
    stores, cents = {}, {}
    t0 = time.time()
    for z in q_by_charge:
        store = dict(per_charge[z][0])
        x = o.vectorize(store["mz"], store["inten"], store["off"])
        nlist = min(wl["nlist"], max(1, len(x) // 39))
        cent = o.kmeans(x, nlist, seed=4, iters=TRAIN_ITERS)
        assign = o.ivf_assign(x, cent)
        store["_list_off"], store["_list_ids"], store["_list_vecs"] = o.build_lists(x, assign, nlist)
        del x
        stores[z], cents[z] = store, cent
    log(f"CPU index build: {time.time() - t0:.1f}s")
    n_all = sum(len(q["prec_mz"]) for q in q_by_charge.values())
    sample = max(32, wl["cpu_sample"] // 4)

    def step(i):
        n = 0
        for z, q in q_by_charge.items():
            take = max(4, int(round(sample * len(q["prec_mz"]) / n_all)))
            lo = (i * take) % max(1, len(q["prec_mz"]) - take)
            qs = synth.take_spectra(q, np.arange(lo, lo + take))
            cpu_pipeline(o, stores[z], qs, cents[z], None, len(cents[z]), wl, z, threads, use_ref)
            n += take
        return n

    for i in range(args.warmup):
        step(i)
    t0 = time.perf_counter()
    n = sum(step(args.warmup + i) for i in range(args.steps))
    dt = time.perf_counter() - t0
    v = round(n / dt, 2)
    desc = {"value": v, "unit": "spectra/s", "cores": threads, "kind": "reference" if use_ref else "port",
            "sample": f"{n // args.steps} queries per step (bounded sample of the {wl['nq']}-query batch)"}
    _emit(json.dumps({
        "impl": "reference", "metric": "query spectra/sec, cascade open search", "value": v, "unit": "spectra/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt / args.steps * 1e3, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32/f64", "data": "synthetic",
        "config": {"workload": wl["label"], "nlist": wl["nlist"], "nprobe": wl["nprobe"], "k": wl["k"],
                   "open_window": f"{OPEN_TOL} {OPEN_MODE}", "fragment_tol": FRAG_TOL},
        "cpu_baseline": desc,
        "e2e": {"value": v, "unit": "spectra/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="solo", choices=["solo", "reference"])
    ap.add_argument("--workload", default=os.environ.get("SOLO_BENCH_WORKLOAD", "c2"), choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--extras", action="store_true",
                    help="also time the widened rows (K6 SSM feature table, .idxann write/read) on the same data; "
                         "adds an 'extras' object to the JSON line")
    ap.add_argument("--sharded", action="store_true",
                    help="mode B: inverted lists sharded over the ranks, one global query batch, NCCL all-gather "
                         "of the per-GPU top-k rows + device merge (strong scaling); default is mode A")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    wl = WORKLOADS[args.workload]
    # stdout carries exactly one JSON line: libraries that print there (NCCL's version banner) go to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    global _emit
    def _emit(line):
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(line, flush=True)
        os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args, wl, rank, world)
    elif "stream_queries" in wl:
        run_stream(args, wl, rank, world, local_rank)
    else:
        run_solo(args, wl, rank, world, local_rank)


if __name__ == "__main__":
    main()
