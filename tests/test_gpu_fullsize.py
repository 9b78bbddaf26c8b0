"""Parity at BASELINE.json's full size (configs[1], "C2": 16,384 queries vs a 3 M-vector library,
nlist 16,384, nprobe 1,024, k 1,024), where the CPU oracle cannot cover the whole batch in seconds:

* size-independent properties of the IVF top-k (sorted by (score desc, id asc), unique ids, every
  score the exact fp32 fmaf dot product of its row — recomputed by the oracle for a sample);
* the tensor-core engine against the exact CUDA-core engine on a slice of the batch (bit-exact ids
  and scores);
* the oracle itself on the smallest charge (300 k vectors) for a handful of queries;
* the fused search: idempotent, winners inside the precursor window, valid, among the top-k.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N_TARGETS, NQ, NLIST, NPROBE, K = 2_000_000, 16384, 16384, 1024, 1024


@pytest.fixture(scope="module")
def c2(synth):
    from ann_solo_b200.engine import SoloEngine
    lib = synth.make_library(N_TARGETS, decoy_fraction=0.5, seed=1, decoy_seed=2)
    per_charge = synth.split_by_charge(lib)
    queries = synth.make_queries(lib, NQ, seed=3)
    eng = SoloEngine(0)
    nlists = {}
    for z in (2, 4):  # the largest and the smallest charge
        store, _ = per_charge[z]
        eng.load_library(z, store)
        nlists[z] = min(NLIST, max(1, len(store["prec_mz"]) // 39))
        eng.ivf_train_library(z, nlists[z], iters=2, seed=4)
        eng.ivf_add_library(z)
    yield eng, per_charge, queries, nlists
    eng.close()


def _queries_of(synth, queries, z, n=None):
    sel = np.flatnonzero(queries["prec_z"] == z)
    return synth.take_spectra(queries, sel[:n] if n else sel)


def test_topk_properties_and_engine_cross_check(c2, oracle, synth):
    eng, per_charge, queries, nlists = c2
    z = 2
    store, _ = per_charge[z]
    q = _queries_of(synth, queries, z, 1024)
    qv = eng.vectorize(q["mz"], q["inten"], q["off"])
    D, I = eng.ivf_search(z, qv, K, NPROBE)
    assert (I >= 0).all() and (I < len(store["prec_mz"])).all()          # ~94 k scanned per query: never short
    assert (np.diff(D, axis=1) <= 0).all()                                # sorted by score
    ties = np.diff(D, axis=1) == 0
    assert (np.diff(I, axis=1)[ties] > 0).all()                           # ties by ascending id
    srt = np.sort(I, axis=1)
    assert (np.diff(srt, axis=1) > 0).all()                               # unique ids
    # every returned score is the oracle's exact dot product of that row (sample of 64 queries x 32 ranks)
    rng = np.random.default_rng(0)
    qs = rng.choice(len(qv), 64, replace=False)
    ranks = rng.choice(K, 32, replace=False)
    rows = np.unique(I[np.ix_(qs, ranks)])
    sub = synth.take_spectra(store, rows)
    xv = oracle.vectorize(sub["mz"], sub["inten"], sub["off"])
    pos = {r: i for i, r in enumerate(rows.tolist())}
    for a in qs:
        for b in ranks:
            assert D[a, b].view(np.uint32) == np.float32(oracle.ip(qv[a], xv[pos[int(I[a, b])]])).view(np.uint32)
    # the exact CUDA-core engine gives the same rows and scores bit for bit
    eng.set_option("scan_engine", 1)
    try:
        D1, I1 = eng.ivf_search(z, qv[:256], K, NPROBE)
    finally:
        eng.set_option("scan_engine", 0)
    assert np.array_equal(I1, I[:256])
    assert np.array_equal(D1.view(np.uint32), D[:256].view(np.uint32))


def test_oracle_parity_on_the_smallest_charge(c2, oracle, synth):
    eng, per_charge, queries, nlists = c2
    z = 4
    store, _ = per_charge[z]
    q = _queries_of(synth, queries, z, 48)
    x = oracle.vectorize(store["mz"], store["inten"], store["off"])
    cent = eng.ivf_get_centroids(z)
    assign = eng.ivf_assignment(z)
    off, ids, vecs = oracle.build_lists(x, assign, len(cent))
    qv = oracle.vectorize(q["mz"], q["inten"], q["off"])
    assert np.array_equal(qv, eng.vectorize(q["mz"], q["inten"], q["off"]))
    Do, Io = oracle.ivf_search(qv, cent, off, ids, vecs, min(NPROBE, len(cent)), K)
    D, I = eng.ivf_search(z, qv, K, NPROBE)
    assert np.array_equal(I, Io)
    assert np.array_equal(D.view(np.uint32), Do.view(np.uint32))


def test_fused_search_properties(c2, synth):
    from ann_solo_b200.engine import SoloEngine
    eng, per_charge, queries, nlists = c2
    z = 2
    store, _ = per_charge[z]
    q = _queries_of(synth, queries, z)
    p = SoloEngine.make_params(True, K, NPROBE, 500.0, "Da", 0.02, True, max_pairs=50)
    r1 = {k: v.copy() for k, v in eng.search_batch(z, p, q).items()}
    r2 = eng.search_batch(z, p, q)
    for key in ("best_row", "score", "n_pairs", "n_cand"):
        assert np.array_equal(r1[key], r2[key]), key                      # idempotent
    has = r1["best_row"] >= 0
    assert has.mean() > 0.99
    assert (r1["n_cand"] <= K).all() and (r1["n_cand"][has] > 0).all()
    rows = r1["best_row"][has]
    lm = store["prec_mz"].astype(np.float32).astype(np.float64)[rows]
    assert (np.abs(q["prec_mz"][has] - lm) * z <= 500.0).all()            # inside the precursor window
    assert store["valid"][rows].all()
    assert ((r1["score"][has] >= 0) & (r1["score"][has] <= 1.0 + 1e-6)).all()  # unit-norm spectra, one-to-one matches
    assert np.array_equal(r1["score"][has] > 0, r1["n_pairs"][has] > 0)       # a score comes from matched peak pairs
    # the winner is one of the query's top-k rows
    qv = eng.vectorize(q["mz"][: q["off"][64]], q["inten"][: q["off"][64]], q["off"][:65])
    _, I = eng.ivf_search(z, qv, K, NPROBE)
    for i in range(64):
        if has[i]:
            assert r1["best_row"][i] in I[i]
    # a query derived from a library spectrum finds that spectrum's precursor neighbourhood: scores well above noise
    assert np.median(r1["score"][has]) > 0.1
