"""K2/K3/K4 (CUDA IVF-Flat) vs the oracle: list assignment, probe sets, top-k ids AND exact fp32
scores bit-for-bit, ties, padding, NaN rows, training."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
CH = 50


def _setup(engine, oracle, synth, n=4000, nlist=32, seed=61, slot=CH):
    lib = synth.make_library(n, decoy_fraction=0.25, seed=seed, decoy_seed=seed + 1)
    x = oracle.vectorize(lib["mz"], lib["inten"], lib["off"])
    cent = oracle.kmeans(x, nlist, seed=4, iters=3)
    engine.ivf_set_centroids(slot, cent)
    engine.ivf_add(slot, x)
    assign = oracle.ivf_assign(x, cent)
    return lib, x, cent, assign


def test_assignment_and_info(engine, oracle, synth):
    lib, x, cent, assign = _setup(engine, oracle, synth)
    assert engine.ivf_info(CH) == (len(x), 32, 800)
    assert np.array_equal(engine.ivf_assignment(CH), assign)
    assert np.array_equal(engine.ivf_get_centroids(CH), cent)


@pytest.mark.parametrize("nprobe,k", [(1, 1), (4, 10), (8, 100), (32, 1024), (100, 50)])
def test_search_bit_exact(engine, oracle, synth, nprobe, k):
    lib, x, cent, assign = _setup(engine, oracle, synth)
    q = synth.make_queries(lib, 200, seed=63)
    qv = oracle.vectorize(q["mz"], q["inten"], q["off"])
    off, ids, vecs = oracle.build_lists(x, assign, 32)
    D, I = engine.ivf_search(CH, qv, k, nprobe)
    Dw, Iw = oracle.ivf_search(qv, cent, off, ids, vecs, nprobe, k)
    assert np.array_equal(I, Iw)
    assert np.array_equal(D, Dw)
    probes = engine.ivf_coarse(CH, qv, nprobe)
    pw, _ = oracle.ivf_coarse(qv, cent, nprobe)
    assert np.array_equal(probes, pw)


def test_incremental_add_equals_single_add(engine, oracle, synth):
    lib, x, cent, assign = _setup(engine, oracle, synth, n=1500, nlist=16, slot=CH + 1)
    engine.ivf_reset(CH + 1)
    assert engine.ivf_info(CH + 1)[0] == 0
    engine.ivf_add(CH + 1, x[:500])
    engine.ivf_add(CH + 1, x[500:501])
    engine.ivf_add(CH + 1, x[501:])
    assert np.array_equal(engine.ivf_assignment(CH + 1), assign)
    off, ids, vecs = oracle.build_lists(x, assign, 16)
    D, I = engine.ivf_search(CH + 1, x[:50], 20, 4)
    Dw, Iw = oracle.ivf_search(x[:50], cent, off, ids, vecs, 4, 20)
    assert np.array_equal(I, Iw) and np.array_equal(D, Dw)


def test_ties_padding_nan(engine, oracle, synth):
    lib = synth.make_library(600, decoy_fraction=0, seed=71)
    x = oracle.vectorize(lib["mz"], lib["inten"], lib["off"])
    x[400] = x[10]          # duplicate vector -> tie broken by id
    x[401] = x[10]
    x[7] = np.nan           # invalid library spectrum (0/0) is not stored
    x[8] = 0                # all-zero row scores 0 everywhere -> list 0
    cent = oracle.kmeans(x, 4, iters=2)
    cent[3] = cent[1]       # duplicate centroid: never chosen over the lower id
    engine.ivf_set_centroids(CH + 2, cent)
    engine.ivf_add(CH + 2, x)
    assign = oracle.ivf_assign(x, cent)
    assert np.array_equal(engine.ivf_assignment(CH + 2), assign)
    assert assign[7] == -1 and assign[8] == 0 and not (assign == 3).any()
    off, ids, vecs = oracle.build_lists(x, assign, 4)
    q = np.ascontiguousarray(x[[10, 20, 8]])
    for nprobe, k in [(4, 5), (2, 700), (4, 2048)]:
        D, I = engine.ivf_search(CH + 2, q, k, nprobe)
        Dw, Iw = oracle.ivf_search(q, cent, off, ids, vecs, nprobe, k)
        assert np.array_equal(I, Iw) and np.array_equal(D, Dw)
    D, I = engine.ivf_search(CH + 2, q, 5, 4)
    assert I[0, :3].tolist() == [10, 400, 401]
    # the zero query scores 0 against everything: pure id order among ties
    assert I[2].tolist() == sorted(I[2].tolist())


def test_large_nlist_many_probes(engine, oracle, synth):
    lib, x, cent, assign = _setup(engine, oracle, synth, n=12000, nlist=1024, seed=81, slot=CH + 3)
    q = synth.make_queries(lib, 64, seed=83)
    qv = oracle.vectorize(q["mz"], q["inten"], q["off"])
    off, ids, vecs = oracle.build_lists(x, assign, 1024)
    D, I = engine.ivf_search(CH + 3, qv, 256, 512)
    Dw, Iw = oracle.ivf_search(qv, cent, off, ids, vecs, 512, 256)
    assert np.array_equal(I, Iw) and np.array_equal(D, Dw)


def test_train_on_device(engine, oracle, synth):
    lib = synth.make_library(3000, decoy_fraction=0, seed=91)
    x = oracle.vectorize(lib["mz"], lib["inten"], lib["off"])
    engine.ivf_train(CH + 4, x, 24, iters=5, seed=7)
    cent = engine.ivf_get_centroids(CH + 4)
    assert cent.shape == (24, 800) and np.isfinite(cent).all()
    np.testing.assert_allclose(np.linalg.norm(cent.astype(np.float64), axis=1), 1.0, atol=1e-5)
    assert engine.ivf_info(CH + 4)[0] == 0  # train leaves no vectors behind (Faiss semantics)
    engine.ivf_add(CH + 4, x)
    assign = oracle.ivf_assign(x, cent)
    assert np.array_equal(engine.ivf_assignment(CH + 4), assign)
    assert np.bincount(assign, minlength=24).min() > 0
    # clustering quality: mean best-centroid similarity clearly above a random pick of rows
    rnd = oracle.ivf_assign(x, x[:24])
    sim_t = np.mean([oracle.ip(x[i], cent[assign[i]]) for i in range(0, 3000, 10)])
    sim_r = np.mean([oracle.ip(x[i], x[rnd[i]]) for i in range(0, 3000, 10)])
    assert sim_t > sim_r


def test_faiss_like_index_api(engine, oracle, synth):
    from ann_solo_b200.index import IndexFlatIP, IndexIVFFlat, METRIC_INNER_PRODUCT
    lib = synth.make_library(1000, decoy_fraction=0, seed=95)
    x = oracle.vectorize(lib["mz"], lib["inten"], lib["off"])
    index = IndexIVFFlat(IndexFlatIP(800), 800, 8, METRIC_INNER_PRODUCT, engine=engine)
    index.train(x)
    index.add(x)
    assert index.ntotal == 1000
    index.nprobe = 8
    D, I = index.search(x[:10], 3)
    assert I.dtype == np.int64 and D.dtype == np.float32 and (I[:, 0] == np.arange(10)).all()
    index.reset()
    assert index.ntotal == 0


def test_compact_probe_selection_equals_dense(engine, oracle, synth):
    """K2's compact path (sampled per-query threshold, thresholded coarse pass, selection over the surviving pairs)
    against the dense (Q, nlist) path and the oracle: identical top-k ids and score bits. Zero and duplicated query
    vectors force the device-side fall-back (threshold 0 -> every list passes -> dense rows for that query)."""
    slot, nlist, nprobe, k = CH + 7, 2048, 128, 64
    lib = synth.make_library(30000, decoy_fraction=0.25, seed=91, decoy_seed=92)
    x = oracle.vectorize(lib["mz"], lib["inten"], lib["off"])
    cent = oracle.kmeans(x, nlist, seed=4, iters=2)
    engine.ivf_set_centroids(slot, cent)
    engine.ivf_add(slot, x)
    assign = oracle.ivf_assign(x, cent)
    off, ids, vecs = oracle.build_lists(x, assign, nlist)
    q = synth.make_queries(lib, 400, seed=93)
    qv = oracle.vectorize(q["mz"], q["inten"], q["off"])
    qv[5] = 0.0                      # all coarse scores 0: the sampled threshold passes every list
    qv[6] = cent[17]                 # a centroid itself
    qv[7, :] = 0.0
    qv[7, 3] = 1.0                   # one non-zero dimension: most scores are exactly 0 (ties at the threshold)
    try:
        engine.set_option("compact_probes", 1)
        D1, I1 = engine.ivf_search(slot, qv, k, nprobe)
        engine.set_option("compact_probes", 0)
        D0, I0 = engine.ivf_search(slot, qv, k, nprobe)
    finally:
        engine.set_option("compact_probes", 1)
    assert np.array_equal(I1, I0) and np.array_equal(D1.view(np.uint32), D0.view(np.uint32))
    Dw, Iw = oracle.ivf_search(qv, cent, off, ids, vecs, nprobe, k)
    assert np.array_equal(I1, Iw) and np.array_equal(D1.view(np.uint32), Dw.view(np.uint32))
    engine.ivf_reset(slot)
