"""Worker of tests/test_gpu_modeb_nccl.py (launched by torch.distributed.run, one process per GPU): mode B over
NCCL — probe rows all-gathered, top-k rows exchanged with all-to-all — must give every rank's query slice exactly
the single-GPU result (ids, score bits, PSMs)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from ann_solo_b200 import parallel, synth
    from ann_solo_b200.engine import SoloEngine
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    charge, nlist, nprobe, k = 2, 96, 32, 128
    lib = synth.make_library(8000, seed=11, decoy_seed=12)
    store, _ = synth.split_by_charge(lib)[charge]
    queries = synth.make_queries(lib, 700, seed=13)
    q = synth.take_spectra(queries, np.flatnonzero(queries["prec_z"] == charge))
    nq = len(q["prec_mz"])
    eng = SoloEngine(local)
    eng.set_stream(torch.cuda.current_stream().cuda_stream)
    eng.load_library(charge, store)
    eng.ivf_train_library(charge, nlist, iters=3, seed=4)
    eng.ivf_add_library(charge)
    params = SoloEngine.make_params(True, k, nprobe, 500.0, "Da", 0.02, True, max_pairs=64)
    ref = {key: v.copy() for key, v in eng.search_batch(charge, params, q).items()}
    assign = eng.ivf_assignment(charge)
    owner = parallel.assign_lists(np.bincount(assign[assign >= 0], minlength=nlist), world)
    eng.ivf_set_owned_lists(charge, (owner == rank).astype(np.uint8))
    bad = 0
    for rep in range(2):   # twice: buffers and streams are reused across batches
        stats = {}
        res = parallel.search_batch_sharded(eng, charge, params, q, rank, world, stats=stats)
        b, e, _ = parallel.slice_bounds(nq, rank, world)
        for key in ("best_row", "n_pairs", "n_cand"):
            bad += int(not np.array_equal(res[key], ref[key][b:e]))
        bad += int(not np.array_equal(res["score"].view(np.uint64), ref["score"][b:e].view(np.uint64)))
        for i in range(e - b):
            n = int(ref["n_pairs"][b + i])
            bad += int(not np.array_equal(res["pairs"][i, :n], ref["pairs"][b + i, :n]))
        assert stats["bytes_sent_per_rank"] > 0
    t = torch.tensor([bad], dtype=torch.int64, device="cuda")
    dist.all_reduce(t)
    if rank == 0:
        print(f"mode B over NCCL, {world} ranks, {nq} queries: {int(t.item())} mismatching fields")
    eng.close()
    dist.destroy_process_group()
    sys.exit(1 if int(t.item()) else 0)


if __name__ == "__main__":
    main()
