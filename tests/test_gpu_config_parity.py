"""Config-scale parity against the CPU oracle (PSM identity, not just properties).

* C1 shape (BASELINE.json configs[0]): 200 k targets + 100 k decoys, every charge, nlist 256, nprobe 128,
  k 1,024, 500 Da open window — >= 512 queries through the fused device path vs the oracle pipeline
  (vectorise -> IVF-Flat (sequential-fmaf definition) -> post-top-k window -> the reference's own compiled
  SpectrumMatch.cpp when oracle/_ref is there): candidate count, best library row, score bits, number of
  matched peaks and the pair set per query (reference: spectral_library.py:328-370, 372-455).
* C2 shape (configs[1]): 3 M-vector library, nlist 16,384, nprobe 1,024, k 1,024 — the same comparison on a
  sample of charge-2 and charge-3 queries (>= 256 each).
* An anchor for the IVF top-k that does not involve the oracle: float64 brute force over the probed lists,
  tie-aware at the probe and top-k boundaries.
"""
import numpy as np
import pytest

from conftest import canon_pairs

pytestmark = pytest.mark.gpu

OPEN_TOL, FRAG_TOL = 500.0, 0.02


def _oracle_psms(oracle, store, q, cent, assign, charge, nprobe, k):
    x = oracle.vectorize(store["mz"], store["inten"], store["off"])
    off, ids, vecs = oracle.build_lists(x, assign, len(cent))
    del x
    qv = oracle.vectorize(q["mz"], q["inten"], q["off"])
    _, ann = oracle.ivf_search(qv, cent, off, ids, vecs, min(nprobe, len(cent)), k)
    del vecs
    cand, coff = oracle.candidates(q["prec_mz"], store["prec_mz"].astype(np.float32), store["valid"], charge, OPEN_TOL,
                                   "Da", ann)
    if oracle.have_ref():  # the reference's own scorer, compiled from where it lies (oracle/Makefile)
        bp, bs, npairs, pairs = oracle.ref_best_match_batch(q, store, cand, coff, FRAG_TOL, True,
                                                            n_threads=oracle.num_threads())
    else:
        bp, bs, npairs, pairs = oracle.best_match_batch(q, store, cand, coff, FRAG_TOL, True, sort_mode=0)
    row = np.full(len(bp), -1, np.int32)
    has = bp >= 0
    row[has] = cand[coff[:-1][has] + bp[has]]
    return row, bs, npairs, pairs, np.diff(coff)


def _assert_same_psms(res, want, label):
    row, bs, npairs, pairs, ncand = want
    n = len(row)
    assert np.array_equal(res["n_cand"][:n], ncand), f"{label}: candidate counts"
    assert np.array_equal(res["best_row"][:n], row), f"{label}: PSM identity"
    has = row >= 0
    assert has.sum() > 0.9 * n, f"{label}: too few matched queries to mean anything"
    assert np.array_equal(res["score"][:n][has].view(np.uint64), bs[has].view(np.uint64)), f"{label}: score bits"
    assert np.array_equal(res["n_pairs"][:n][has], npairs[has]), f"{label}: matched peak counts"
    for i in np.flatnonzero(has):
        m = int(npairs[i])
        assert np.array_equal(canon_pairs(res["pairs"][i], m), canon_pairs(pairs[i], m)), f"{label}: pairs of query {i}"


def test_c1_shape_full_pipeline_equals_oracle(oracle, synth):
    from ann_solo_b200.engine import SoloEngine
    lib = synth.make_library(200_000, decoy_fraction=0.5, seed=1, decoy_seed=2)
    per_charge = synth.split_by_charge(lib)
    queries = synth.make_queries(lib, 640, seed=3)
    eng = SoloEngine(0)
    try:
        p = SoloEngine.make_params(True, 1024, 128, OPEN_TOL, "Da", FRAG_TOL, True, max_pairs=50)
        checked = 0
        for z in sorted(per_charge):
            store, _ = per_charge[z]
            q = synth.take_spectra(queries, np.flatnonzero(queries["prec_z"] == z))
            if len(q["prec_mz"]) == 0:
                continue
            nlist = min(256, max(1, len(store["prec_mz"]) // 39))
            eng.load_library(z, store)
            eng.ivf_train_library(z, nlist, iters=2, seed=4)
            eng.ivf_add_library(z)
            res = eng.search_batch(z, p, q)
            want = _oracle_psms(oracle, store, q, eng.ivf_get_centroids(z), eng.ivf_assignment(z), z, 128, 1024)
            _assert_same_psms(res, want, f"C1 charge {z}")
            checked += len(q["prec_mz"])
        assert checked >= 512
    finally:
        eng.close()


@pytest.fixture(scope="module")
def c2_world(synth):
    lib = synth.make_library(2_000_000, decoy_fraction=0.5, seed=1, decoy_seed=2)
    per_charge = synth.split_by_charge(lib)
    queries = synth.make_queries(lib, 16384, seed=3)
    return per_charge, queries


@pytest.mark.parametrize("z", [2, 3])
def test_c2_sample_equals_oracle(c2_world, oracle, synth, z):
    from ann_solo_b200.engine import SoloEngine
    per_charge, queries = c2_world
    store, _ = per_charge[z]
    q = synth.take_spectra(queries, np.flatnonzero(queries["prec_z"] == z)[:288])
    assert len(q["prec_mz"]) >= 256
    eng = SoloEngine(0)
    try:
        nlist = min(16384, max(1, len(store["prec_mz"]) // 39))
        eng.load_library(z, store)
        eng.ivf_train_library(z, nlist, iters=2, seed=4)
        eng.ivf_add_library(z)
        p = SoloEngine.make_params(True, 1024, 1024, OPEN_TOL, "Da", FRAG_TOL, True, max_pairs=50)
        res = eng.search_batch(z, p, q)
        want = _oracle_psms(oracle, store, q, eng.ivf_get_centroids(z), eng.ivf_assignment(z), z, 1024, 1024)
        _assert_same_psms(res, want, f"C2 charge {z}")
    finally:
        eng.close()


def test_ivf_topk_against_float64_brute_force(engine, synth):
    """Independent anchor: inner products in float64 (NumPy), no oracle code. A query is compared when its
    probe boundary is unambiguous in float64 (gap between the nprobe-th and the next centroid score > 1e-6);
    the id sets must then agree except inside the tie band around the k-th score (|s - s_k| <= 2e-6, the
    float32 rounding of a unit-norm dot product)."""
    lib = synth.make_library(40_000, seed=11, decoy_seed=12)
    store, _ = synth.split_by_charge(lib)[2]
    queries = synth.make_queries(lib, 400, seed=13)
    q = synth.take_spectra(queries, np.flatnonzero(queries["prec_z"] == 2)[:96])
    nlist, nprobe, k = 128, 24, 256
    engine.set_vectorizer(11, 2010, 0.04, 800)
    engine.load_library(2, store)
    engine.ivf_train_library(2, nlist, iters=3, seed=4)
    engine.ivf_add_library(2)
    x = engine.vectorize(store["mz"], store["inten"], store["off"]).astype(np.float64)
    qv32 = engine.vectorize(q["mz"], q["inten"], q["off"])
    qv = qv32.astype(np.float64)
    cent = engine.ivf_get_centroids(2).astype(np.float64)
    assign = engine.ivf_assignment(2)
    D, I = engine.ivf_search(2, qv32, k, nprobe)
    coarse = qv @ cent.T
    compared = 0
    for i in range(len(qv)):
        order = np.argsort(-coarse[i], kind="stable")
        if coarse[i, order[nprobe - 1]] - coarse[i, order[nprobe]] <= 1e-6:
            continue  # ambiguous probe set in float64: not a fair comparison
        rows = np.flatnonzero(np.isin(assign, order[:nprobe]))
        s = x[rows] @ qv[i]
        assert len(rows) >= k
        top = np.argsort(-s, kind="stable")
        s_k = s[top[k - 1]]
        sure_in = set(rows[s > s_k + 2e-6].tolist())
        maybe = set(rows[np.abs(s - s_k) <= 2e-6].tolist())
        got = set(I[i].tolist())
        assert len(got) == k
        assert sure_in <= got, f"query {i}: rows above the tie band are missing"
        assert got <= (sure_in | maybe), f"query {i}: rows below the tie band were returned"
        # and the scores are the float32 view of the float64 ones
        pos = {r: j for j, r in enumerate(rows.tolist())}
        ref = np.array([s[pos[r]] for r in I[i].tolist()])
        assert np.abs(D[i] - ref).max() <= 2e-6
        compared += 1
    assert compared >= 64
