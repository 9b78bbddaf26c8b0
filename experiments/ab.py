"""A/B timing of engine options on the C2 workload (data and index built once).
usage: python experiments/ab.py "scan_ts=0" "scan_ts=96" "scan_ts=112,tc_kbb=2" ...   (each argument = one variant)"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from ann_solo_b200.engine import SoloEngine
wl = bench.WORKLOADS[os.environ.get("AB_WORKLOAD", "c2")]
lib, per_charge, q_by_charge = bench.make_data(wl, 0)
torch.cuda.set_stream(torch.cuda.Stream())
eng = SoloEngine(0)
eng.set_stream(torch.cuda.current_stream().cuda_stream)
charges = sorted(q_by_charge)
for z in charges:
    store, _ = per_charge[z]
    eng.load_library(z, store)
    nlist = min(wl["nlist"], max(1, len(store["prec_mz"]) // 39))
    eng.ivf_train_library(z, nlist, iters=bench.TRAIN_ITERS, seed=4)
    eng.ivf_add_library(z)
params = SoloEngine.make_params(True, wl["k"], wl["nprobe"], bench.OPEN_TOL, bench.OPEN_MODE, bench.FRAG_TOL, True, 50)
for z in charges:
    eng.select_slot(z)
    eng.stage_queries(q_by_charge[z])
eng.synchronize()
def step():
    for z in charges:
        eng.select_slot(z)
        eng.search_staged(z, params)
ref = None
for variant in sys.argv[1:] or ["scan_ts=0"]:
    for kv in variant.split(","):
        k, v = kv.split("=")
        eng.set_option(k, int(v))
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    eng.profile_reset(); eng.profile_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 10
    e0.record()
    for _ in range(steps):
        step()
    e1.record(); torch.cuda.synchronize()
    prof = eng.profile(); eng.profile_enable(False)
    ms = e0.elapsed_time(e1) / steps
    # checksum of the results for equality between variants
    sig = []
    for z in charges:
        eng.select_slot(z)
        r = eng.fetch_results()
        sig.append((int(r["best_row"].astype(np.int64).sum()), float(r["score"].sum()), int(r["n_cand"].sum())))
    same = "" if ref is None else (" results==first" if sig == ref else " RESULTS DIFFER")
    ref = ref or sig
    st = {k: round(v["ms"] / steps, 3) for k, v in prof.items() if v["ms"] > 0}
    print(f"{variant:40s} {ms:8.3f} ms/step  {st}{same}", flush=True)
