"""Dump the C2 index shape (list lengths, queries per list) to gpurun_out/c2_shape.npz for off-line tile-count models."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from ann_solo_b200.engine import SoloEngine
wl = bench.WORKLOADS["c2"]
lib, per_charge, q_by_charge = bench.make_data(wl, 0)
eng = SoloEngine(0)
bal = int(os.environ.get("BALANCE", "0"))
if bal:
    eng.set_option("train_balance", bal)
out = {}
for z in sorted(q_by_charge):
    store, _ = per_charge[z]
    eng.load_library(z, store)
    nlist = min(wl["nlist"], max(1, len(store["prec_mz"]) // 39))
    t0 = time.time()
    eng.ivf_train_library(z, nlist, iters=int(os.environ.get("SOLO_TRAIN_ITERS", "2")), seed=4)
    eng.ivf_add_library(z)
    eng.synchronize()
    a = eng.ivf_assignment(z)
    sizes = np.bincount(a[a >= 0], minlength=nlist)
    q = q_by_charge[z]
    qv = eng.vectorize(q["mz"], q["inten"], q["off"])
    probes = eng.ivf_coarse(z, qv, min(wl["nprobe"], nlist))
    G = np.bincount(probes.ravel(), minlength=nlist)
    out[f"sizes{z}"] = sizes.astype(np.int32)
    out[f"G{z}"] = G.astype(np.int32)
    print(z, nlist, "train", round(time.time() - t0, 1), "s; sizes mean", sizes.mean(), "median", np.median(sizes), "max", sizes.max(),
          "G mean", G.mean(), "corr", np.corrcoef(sizes, G)[0, 1], flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.savez_compressed(os.path.join(ROOT, "gpurun_out", f"c2_shape_b{bal}.npz"), **out)
