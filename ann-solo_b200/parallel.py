"""Multi-GPU plumbing (one process per GPU, torch.distributed; NCCL on GPUs, gloo in CPU tests).

Mode A (library fits in HBM): the library and index are replicated, the queries of
every batch are partitioned contiguously by rank, there is NO collective on the data path; the
per-rank results are gathered to rank 0 once.

Mode B (IVF lists sharded): every rank scans only the lists it owns; the per-rank top-k
(score, id) rows are exchanged with one all-gather and merged under the same total order
(score desc, id asc), so the merged top-k equals the single-GPU top-k exactly.
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np


def shard_bounds(n: int, rank: int, world: int):
    """Contiguous, balanced [begin, end) of `n` items for `rank` of `world`."""
    base, rem = divmod(n, world)
    b = rank * base + min(rank, rem)
    return b, b + base + (1 if rank < rem else 0)


def shard_store(store: dict, rank: int, world: int) -> dict:
    """The rank's contiguous slice of a query peak store (CSR re-based to 0)."""
    n = len(store["off"]) - 1
    b, e = shard_bounds(n, rank, world)
    p0, p1 = store["off"][b], store["off"][e]
    out = {}
    for k, v in store.items():
        if k == "off":
            out[k] = (v[b:e + 1] - p0).astype(np.int64)
        elif len(v) == n:
            out[k] = v[b:e]
        else:
            out[k] = v[p0:p1]
    return out


def assign_lists(list_sizes: np.ndarray, world: int) -> np.ndarray:
    """Mode B: owner rank of every inverted list, balancing the number of stored vectors
    (longest-processing-time greedy; deterministic)."""
    order = np.argsort(-np.asarray(list_sizes), kind="stable")
    load = np.zeros(world, np.int64)
    owner = np.empty(len(list_sizes), np.int32)
    for l in order:
        r = int(np.argmin(load))
        owner[l] = r
        load[r] += list_sizes[l]
    return owner


def merge_topk(D_parts: List[np.ndarray], I_parts: List[np.ndarray], k: int):
    """Merge per-rank (nq, k) results under (score desc, id asc); -1 ids are padding."""
    D = np.concatenate(D_parts, axis=1)
    I = np.concatenate(I_parts, axis=1)
    Dk = np.where(I < 0, -np.inf, D)
    Ik = np.where(I < 0, np.iinfo(np.int64).max, I)
    order = np.lexsort((Ik, -Dk), axis=1)[:, :k]
    Dm = np.take_along_axis(D, order, 1)
    Im = np.take_along_axis(I, order, 1)
    Dm = np.where(Im < 0, -np.inf, Dm).astype(np.float32)
    return Dm, Im


def gather_results(res: dict, dst: int = 0) -> Optional[dict]:
    """Concatenate per-rank result dicts (arrays with one row per query) on `dst` in rank order."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return res
    parts = [None] * dist.get_world_size() if dist.get_rank() == dst else None
    dist.gather_object(res, parts, dst=dst)
    if dist.get_rank() != dst:
        return None
    return {k: np.concatenate([p[k] for p in parts], axis=0) for k in res}


def allgather_topk(D: np.ndarray, I: np.ndarray, k: int):
    """Mode B exchange: one all-gather of the per-rank (D, I) rows, then the merge."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return D, I
    world = dist.get_world_size()
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    tD = torch.from_numpy(np.ascontiguousarray(D)).to(dev)
    tI = torch.from_numpy(np.ascontiguousarray(I)).to(dev)
    gD = [torch.empty_like(tD) for _ in range(world)]
    gI = [torch.empty_like(tI) for _ in range(world)]
    dist.all_gather(gD, tD)
    dist.all_gather(gI, tI)
    return merge_topk([t.cpu().numpy() for t in gD], [t.cpu().numpy() for t in gI], k)


def search_batch_sharded(eng, charge: int, params, q: dict, rank: int = 0, world: int = 1, group=None,
                         peers=None) -> dict:
    """Mode B for one batch: every GPU holds the whole query batch and a shard of the inverted
    lists (``SoloEngine.ivf_set_owned_lists``). Each GPU scans its lists (device top-k rows), the
    rows are exchanged with ONE all-gather (NCCL over NVLink), every GPU merges and finishes
    (precursor window, best match) its contiguous slice of the queries on the device. Returns the
    slice's results (rows ``shard_bounds(nq, rank, world)``).

    ``peers``: single-process variant used by the one-GPU test — a list of engines standing in for
    the ranks (each owning some lists); the exchange is a concatenation instead of a collective.
    """
    import torch
    dev = torch.device("cuda", eng.device)
    nq = len(q["off"]) - 1
    k = params.k
    engines = peers if peers is not None else [eng]
    parts_D, parts_I = [], []
    for e in engines:
        e.stage_queries(q)
        I = torch.empty((nq, k), dtype=torch.int64, device=dev)
        D = torch.empty((nq, k), dtype=torch.float32, device=dev)
        e.ivf_search_staged(charge, k, params.nprobe, I.data_ptr(), D.data_ptr())
        e.synchronize()
        parts_D.append(D)
        parts_I.append(I)
    if peers is not None:
        all_D, all_I, parts = torch.stack(parts_D), torch.stack(parts_I), len(peers)
    elif world > 1:
        import torch.distributed as dist
        all_D = torch.empty((world, nq, k), dtype=torch.float32, device=dev)
        all_I = torch.empty((world, nq, k), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(all_D, parts_D[0], group=group)
        dist.all_gather_into_tensor(all_I, parts_I[0], group=group)
        torch.cuda.synchronize(dev)
        parts = world
    else:
        all_D, all_I, parts = parts_D[0][None], parts_I[0][None], 1
    b, e_ = shard_bounds(nq, rank, world)
    mD = torch.empty((e_ - b, k), dtype=torch.float32, device=dev)
    mI = torch.empty((e_ - b, k), dtype=torch.int64, device=dev)
    eng.merge_topk_device(all_D.data_ptr(), all_I.data_ptr(), parts, nq, k, b, e_ - b, mD.data_ptr(), mI.data_ptr())
    eng.score_staged_ids(charge, params, mI.data_ptr(), b, e_ - b)
    res = eng.fetch_results()
    return {key: v[b:e_] for key, v in res.items()}, (mD, mI)
