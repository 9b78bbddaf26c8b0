"""How often are neighbouring inverted lists probed by the same queries? (C2, per charge)"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from ann_solo_b200.engine import SoloEngine
wl = bench.WORKLOADS["c2"]
lib, per_charge, q_by_charge = bench.make_data(wl, 0)
eng = SoloEngine(0)
for z in sorted(q_by_charge):
    store, _ = per_charge[z]
    eng.load_library(z, store)
    nlist = min(wl["nlist"], max(1, len(store["prec_mz"]) // 39))
    eng.ivf_train_library(z, nlist, iters=2, seed=4)
    eng.ivf_add_library(z)
    a = eng.ivf_assignment(z)
    sizes = torch.from_numpy(np.bincount(a[a >= 0], minlength=nlist)).cuda()
    q = q_by_charge[z]
    qv = eng.vectorize(q["mz"], q["inten"], q["off"])
    probes = torch.from_numpy(eng.ivf_coarse(z, qv, min(wl["nprobe"], nlist)).astype(np.int64)).cuda()
    Q = probes.shape[0]
    M = torch.zeros((Q, nlist), dtype=torch.float16, device="cuda")
    M.scatter_(1, probes, 1.0)
    cent = torch.from_numpy(eng.ivf_get_centroids(z)).cuda()
    sim = cent @ cent.T
    sim.fill_diagonal_(-1)
    # co-probe counts between every pair of lists: (nlist, nlist) = M^T M
    co = (M.T @ M).float()
    G = co.diagonal().clone()
    co.fill_diagonal_(0)
    for name, partner in (("nearest centroid", sim.argmax(1)), ("most co-probed", co.argmax(1))):
        inter = co[torch.arange(nlist, device="cuda"), partner]
        union = G + G[partner] - inter
        J = (inter / union.clamp(min=1)).cpu().numpy()
        short = (sizes <= 48).cpu().numpy()
        print(f"z={z} partner={name}: Jaccard mean {J.mean():.3f} median {np.median(J):.3f} p90 {np.percentile(J, 90):.3f}; "
              f"short lists (<=48): mean {J[short].mean():.3f}; union/sum mean {(union / (G + G[partner]).clamp(min=1)).mean().item():.3f}", flush=True)
    # greedy pairing of short lists with their most co-probed short partner
    del sim
