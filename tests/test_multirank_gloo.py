"""world_size-2 gloo runs of the multi-GPU host logic (mode A query partition + gather, mode B
top-k all-gather merge) on CPU; the per-rank compute is the oracle here because this test is
about the partition/exchange code, which is backend-agnostic."""
import os
import sys

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ann_solo_b200 import parallel, synth
    from oracle import solo_oracle as o
    lib = synth.make_library(800, seed=1, decoy_seed=2)
    q = synth.make_queries(lib, 41, seed=3)
    # ---- mode A: partition queries, replicate library, gather
    mine = parallel.shard_store(q, rank, world)
    nq = len(mine["prec_mz"])
    rng = np.random.default_rng(5)
    cand_all = np.sort(rng.integers(0, len(lib["prec_mz"]), (41, 16)), axis=1).astype(np.int32)
    b, e = parallel.shard_bounds(41, rank, world)
    bp, bs, npairs, pairs = o.best_match_batch(mine, lib, cand_all[b:e].ravel(),
                                               np.arange(0, nq * 16 + 1, 16, dtype=np.int64), 0.02, True)
    res = parallel.gather_results(dict(best=bp, score=bs, n_pairs=npairs), dst=0)
    # ---- mode B: lists sharded, all-gather merge
    x = o.vectorize(lib["mz"], lib["inten"], lib["off"])
    cent = o.kmeans(x, 8, iters=2)
    assign = o.ivf_assign(x, cent)
    full_off, _, _ = o.build_lists(x, assign, 8)
    owner = parallel.assign_lists(np.diff(full_off), world)
    a_r = np.where(owner[np.maximum(assign, 0)] == rank, assign, -1)
    o_r, i_r, v_r = o.build_lists(x, a_r, 8)
    qv = o.vectorize(q["mz"], q["inten"], q["off"])
    d, i = o.ivf_search(qv, cent, o_r, i_r, v_r, nprobe=4, k=30)
    Dm, Im = parallel.allgather_topk(d, i, 30)
    if rank == 0:
        np.savez(tmp, best=res["best"], score=res["score"], n_pairs=res["n_pairs"], Dm=Dm, Im=Im)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_partition_and_merge(tmp_path, oracle, synth):
    out = str(tmp_path / "r0.npz")
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    lib = synth.make_library(800, seed=1, decoy_seed=2)
    q = synth.make_queries(lib, 41, seed=3)
    rng = np.random.default_rng(5)
    cand = np.sort(rng.integers(0, len(lib["prec_mz"]), (41, 16)), axis=1).astype(np.int32)
    bp, bs, npairs, _ = oracle.best_match_batch(q, lib, cand.ravel(), np.arange(0, 41 * 16 + 1, 16, dtype=np.int64),
                                                0.02, True)
    assert np.array_equal(got["best"], bp) and np.array_equal(got["score"], bs)
    assert np.array_equal(got["n_pairs"], npairs)
    x = oracle.vectorize(lib["mz"], lib["inten"], lib["off"])
    cent = oracle.kmeans(x, 8, iters=2)
    off, ids, vecs = oracle.build_lists(x, oracle.ivf_assign(x, cent), 8)
    qv = oracle.vectorize(q["mz"], q["inten"], q["off"])
    D, I = oracle.ivf_search(qv, cent, off, ids, vecs, nprobe=4, k=30)
    assert np.array_equal(got["Im"], I) and np.array_equal(got["Dm"], D)
