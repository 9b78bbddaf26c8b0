"""Multi-GPU plumbing (one process per GPU, torch.distributed; NCCL on GPUs, gloo in CPU tests).

Mode A (library fits in HBM): the library and index are replicated, the queries of
every batch are partitioned contiguously by rank, there is NO collective on the data path; the
per-rank results are gathered to rank 0 once.

Mode B (IVF lists sharded): coarse scoring is sharded by queries (probe rows all-gathered), every
rank scans only the lists it owns, the per-rank top-k (score, id) rows travel to the rank that owns the
query slice with one all-to-all and are merged there under the same total order (score desc, id asc),
so the merged top-k equals the single-GPU top-k exactly.
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


_PACKED_PAD = -0x7FFFFF00000001   # 0xFF800000FFFFFFFF as int64: score -inf, row -1 (solo_b200.h)


def slice_bounds(n: int, rank: int, world: int):
    """Mode B query slices: equal length ceil(n / world) (what all_gather_into_tensor / all_to_all_single
    exchange), the last ones clipped at n. Returns (begin, end, slice_len)."""
    s = -(-n // world) if n else 0
    b = min(rank * s, n)
    return b, min(b + s, n), s


def search_batch_sharded(eng, charge: int, params, q: dict, rank: int = 0, world: int = 1, group=None,
                         peers=None, mz_vec: Optional[np.ndarray] = None, stats: Optional[dict] = None):
    """Mode B for one batch: every GPU holds the whole query batch and a shard of the inverted lists
    (``SoloEngine.ivf_set_owned_lists``). Work is split twice:

    1. coarse scoring + probe selection by QUERIES: rank r does it for its slice, the (slice, nprobe) int32 rows
       are all-gathered (NCCL over NVLink) — every GPU then knows every query's lists;
    2. the list scan + exact top-k by LISTS: every GPU scans the lists it owns for all queries;
    3. the per-GPU top-k entries — unsorted, 8 bytes each: (approximate score, library row) — go to the GPU that
       owns the query slice with ONE all-to-all (rank r receives only its slice's rows from every peer:
       Q k 8 / world bytes per peer instead of the whole (Q, k) tensor);
    4. every GPU selects the exact global top-k of its slice with the single-GPU band rule (entries near the k-th
       score are re-scored exactly from the sparse rows every GPU keeps) and finishes it (precursor window, best
       match).

    Everything runs on the engine's stream = torch's current stream, so the collectives are ordered behind the
    kernels without host synchronisation. Returns the results of the slice; the candidate sets, hence the SSMs,
    equal the single-GPU ones exactly.

    ``peers``: single-process variant used by the one-GPU test — a list of engines standing in for the ranks (each
    owning some lists); the exchanges are tensor copies instead of collectives.
    """
    import torch
    dev = torch.device("cuda", eng.device)
    nq = len(q["off"]) - 1
    k, nprobe = params.k, params.nprobe
    engines = peers if peers is not None else [eng]
    parts = len(peers) if peers is not None else world
    S = slice_bounds(nq, 0, parts)[2]
    stream = torch.cuda.current_stream(dev).cuda_stream
    for e in engines:
        e.set_stream(stream)
        e.stage_queries(q, mz_vec)
        # (the unconditional first scan round keeps its single-GPU size: every GPU needs its own k-th best score as the
        # running threshold, and a round of 2 k scores leaves a threshold that lets a third of the later scores pass —
        # measured at 2 GPUs: scan 5.5 -> 7.7 ms, top-k 2.7 -> 4.1 ms)
    # 1. probes, sharded by queries
    probes_all = torch.zeros((parts * S, nprobe), dtype=torch.int32, device=dev)
    if peers is not None:
        for r, e in enumerate(engines):
            b, en, _ = slice_bounds(nq, r, parts)
            e.ivf_probe_staged(charge, nprobe, b, en - b, probes_all[r * S:].data_ptr())
    else:
        b, en, _ = slice_bounds(nq, rank, world)
        if world > 1:
            import torch.distributed as dist
            mine = torch.zeros((S, nprobe), dtype=torch.int32, device=dev)
            eng.ivf_probe_staged(charge, nprobe, b, en - b, mine.data_ptr())
            dist.all_gather_into_tensor(probes_all, mine, group=group)
        else:
            eng.ivf_probe_staged(charge, nprobe, b, en - b, probes_all.data_ptr())
    # 2. scan of the owned lists for every query, 3. exchange towards the owner of each query slice: packed
    #    (score bits << 32 | row) entries, 8 bytes each, unsorted
    recv = torch.empty((parts, S, k), dtype=torch.int64, device=dev)
    my = rank
    for r, e in enumerate(engines):
        loc = torch.empty((parts * S, k), dtype=torch.int64, device=dev)
        if parts * S > nq:
            loc[nq:].fill_(_PACKED_PAD)
        e.ivf_scan_staged(charge, k, nprobe, probes_all.data_ptr(), d_packed=loc.data_ptr())
        if peers is not None:
            recv[r].copy_(loc[my * S:(my + 1) * S])
        elif world > 1:
            import torch.distributed as dist
            dist.all_to_all_single(recv.view(parts * S, k), loc, group=group)
        else:
            recv = loc.view(1, S, k)
    if stats is not None:
        stats["bytes_sent_per_rank"] = int((parts - 1) * S * (nprobe * 4 + k * 8)) if parts > 1 else 0
    # 4. exact global top-k of the slice (band re-scored from the replicated sparse rows), window, best match
    b, en, _ = slice_bounds(nq, my, parts)
    eng.merge_score_staged(charge, params, recv.data_ptr(), parts, S, b, en - b)
    return eng.fetch_results_range(b, en - b)
