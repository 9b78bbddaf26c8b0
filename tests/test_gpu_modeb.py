"""Mode B (inverted lists sharded over GPUs) on ONE GPU: two engines stand in for two
ranks, each storing half of the lists; the device merge of their top-k rows and the sliced finish
must reproduce the single-engine result exactly."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _build(eng, store, cent, charge, owned=None):
    eng.load_library(charge, store)
    eng.ivf_set_centroids(charge, cent)
    if owned is not None:
        eng.ivf_set_owned_lists(charge, owned)
    eng.ivf_add_library(charge)


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_lists_equal_single_gpu(engine, oracle, synth, world):
    import torch
    from ann_solo_b200 import parallel
    from ann_solo_b200.engine import SoloEngine
    charge, nlist, nprobe, k = 2, 48, 24, 96
    lib = synth.make_library(4000, seed=11, decoy_seed=12)
    store, _ = synth.split_by_charge(lib)[charge]
    queries = synth.make_queries(lib, 300, seed=13)
    q = synth.take_spectra(queries, np.flatnonzero(queries["prec_z"] == charge))
    nq = len(q["prec_mz"])
    vec = engine.vectorize(store["mz"], store["inten"], store["off"])
    cent = oracle.kmeans(vec, nlist, seed=4, iters=3)
    params = SoloEngine.make_params(True, k, nprobe, 500.0, "Da", 0.02, True, max_pairs=64)

    _build(engine, store, cent, charge)
    ref = engine.search_batch(charge, params, q)
    D_ref, I_ref = engine.ivf_search(charge, engine.vectorize(q["mz"], q["inten"], q["off"]), k, nprobe)

    sizes = np.bincount(engine.ivf_assignment(charge)[engine.ivf_assignment(charge) >= 0], minlength=nlist)
    owner = parallel.assign_lists(sizes, world)
    peers = []
    for r in range(world):
        e = SoloEngine(0)
        _build(e, store, cent, charge, (owner == r).astype(np.uint8))
        peers.append(e)
    got = {}
    for r in range(world):  # every "rank" merges and finishes its own slice of the queries
        res = parallel.search_batch_sharded(peers[r], charge, params, q, rank=r, world=world, peers=peers)
        for key, v in res.items():
            got.setdefault(key, []).append(v)
    # the sorted (D, I) form of the local search still merges to the single-GPU top-k bit for bit
    dev = torch.device("cuda", 0)
    pr = torch.empty((nq, nprobe), dtype=torch.int32, device=dev)
    peers[0].stage_queries(q)
    peers[0].ivf_probe_staged(charge, nprobe, 0, nq, pr.data_ptr())
    aD = torch.empty((world, nq, k), dtype=torch.float32, device=dev)
    aI = torch.empty((world, nq, k), dtype=torch.int64, device=dev)
    for r in range(world):
        peers[r].stage_queries(q)
        peers[r].ivf_scan_staged(charge, k, nprobe, pr.data_ptr(), aI[r].data_ptr(), aD[r].data_ptr())
        peers[r].synchronize()
    mD = torch.empty((nq, k), dtype=torch.float32, device=dev)
    mI = torch.empty((nq, k), dtype=torch.int64, device=dev)
    peers[0].merge_topk_device(aD.data_ptr(), aI.data_ptr(), world, nq, k, 0, nq, mD.data_ptr(), mI.data_ptr())
    peers[0].synchronize()
    assert np.array_equal(mI.cpu().numpy(), I_ref), "merged ids differ from the single-GPU top-k"
    assert np.array_equal(mD.cpu().numpy().view(np.uint32), D_ref.view(np.uint32))
    for key in ("best_row", "score", "n_pairs", "n_cand"):
        assert np.array_equal(np.concatenate(got[key]), ref[key]), key
    pairs = np.concatenate(got["pairs"])
    for i in range(nq):
        n = ref["n_pairs"][i]
        assert np.array_equal(pairs[i, :n], ref["pairs"][i, :n])
    for e in peers:
        e.close()
    engine.ivf_set_owned_lists(charge, None)
