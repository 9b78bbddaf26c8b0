"""The tcgen05/TMA list-scan engine: raw approximate scores against an fp16-input model, the
proven error bound against the oracle's exact fp32 chain, and bit-exact top-k after the band
re-rank (cross-checked with the exact CUDA-core engine)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
CH = 60
REL_EPS = 1.25e-3  # IVF_REL_EPS in csrc/ivf.cuh


def _index(engine, oracle, synth, n, nlist, seed):
    lib = synth.make_library(n, decoy_fraction=0.0, seed=seed)
    x = oracle.vectorize(lib["mz"], lib["inten"], lib["off"])
    cent = oracle.kmeans(x, nlist, seed=4, iters=2)
    engine.ivf_set_centroids(CH, cent)
    engine.ivf_add(CH, x)
    return lib, x, cent


@pytest.mark.parametrize("n,nlist,nq", [(700, 4, 37), (1900, 4, 300), (2000, 24, 200)])
def test_raw_scores_match_fp16_model(engine, oracle, synth, n, nlist, nq):
    """Every (query, vector) pair is scanned (nprobe = nlist) and dumped before selection."""
    lib, x, cent = _index(engine, oracle, synth, n, nlist, seed=111)
    q = synth.make_queries(lib, nq, seed=113)
    qv = oracle.vectorize(q["mz"], q["inten"], q["off"])
    engine.set_option("scan_engine", 0)
    engine.ivf_search(CH, qv, 2048, nlist)  # k > n: the buffers are not compacted
    counts, dump = engine.debug_scan_dump(CH, nq)
    assert (counts == n).all()  # lists hold everything, nothing filtered in round 0
    qh = (qv * 1024).astype(np.float16).astype(np.float64)
    xh = (x * 1024).astype(np.float16).astype(np.float64)
    model = qh @ xh.T / 2.0 ** 20
    exact = qv.astype(np.float64) @ x.astype(np.float64).T
    worst_model, worst_rel = 0.0, 0.0
    for i in range(nq):
        s, rows = dump[i]
        assert len(np.unique(rows)) == n
        got = np.empty(n)
        got[rows] = s
        worst_model = max(worst_model, np.abs(got - model[i]).max())
        worst_rel = max(worst_rel, (np.abs(got - exact[i]) / np.maximum(exact[i], 1e-3)).max())
    assert worst_model < 2e-6          # fp32 accumulation of exact fp16 products
    assert worst_rel < 1.0e-3 < REL_EPS  # two fp16 roundings (2 * 2^-11) + fp32 accumulation, inside the bound


@pytest.mark.parametrize("n,nlist,nprobe,k,nq", [(3000, 4, 4, 64, 300), (6000, 64, 16, 128, 513),
                                                 (20000, 256, 64, 1024, 257), (1200, 3, 2, 2048, 50)])
def test_topk_bit_exact_both_engines(engine, oracle, synth, n, nlist, nprobe, k, nq):
    lib, x, cent = _index(engine, oracle, synth, n, nlist, seed=121)
    x2 = x.copy()
    q = synth.make_queries(lib, nq, seed=123)
    qv = oracle.vectorize(q["mz"], q["inten"], q["off"])
    off, ids, vecs = oracle.build_lists(x2, oracle.ivf_assign(x2, cent), nlist)
    Dw, Iw = oracle.ivf_search(qv, cent, off, ids, vecs, nprobe, k)
    try:
        for eng_id in (0, 1):
            engine.set_option("scan_engine", eng_id)
            D, I = engine.ivf_search(CH, qv, k, nprobe)
            assert np.array_equal(I, Iw), f"engine {eng_id}"
            assert np.array_equal(D, Dw), f"engine {eng_id}"
        # the swapped-operand scan (list chunk in tensor memory, scan_ts_kernel): 96 and 112 queries per tile
        engine.set_option("scan_engine", 0)
        for nq_tile in (96, 112):
            engine.set_option("scan_ts", nq_tile)
            D, I = engine.ivf_search(CH, qv, k, nprobe)
            assert np.array_equal(I, Iw), f"scan_ts {nq_tile}"
            assert np.array_equal(D, Dw), f"scan_ts {nq_tile}"
    finally:
        engine.set_option("scan_engine", 0)
        engine.set_option("scan_ts", 0)


def test_lists_of_twenty_thousand_vectors(engine, oracle, synth):
    """Inverted lists far beyond the 96-row chunk and beyond the former 12,288-vector cap (C5-sized lists): many chunks
    per list, a first scan round that appends a whole 20 k-vector list, the large-capacity top-k launch."""
    n, nlist, nprobe, k, nq = 60000, 3, 2, 1024, 48
    lib, x, cent = _index(engine, oracle, synth, n, nlist, seed=141)
    q = synth.make_queries(lib, nq, seed=143)
    qv = oracle.vectorize(q["mz"], q["inten"], q["off"])
    assign = oracle.ivf_assign(x, cent)
    assert np.bincount(assign, minlength=nlist).max() > 12288
    off, ids, vecs = oracle.build_lists(x, assign, nlist)
    Dw, Iw = oracle.ivf_search(qv, cent, off, ids, vecs, nprobe, k)
    engine.set_option("scan_engine", 0)
    D, I = engine.ivf_search(CH, qv, k, nprobe)
    assert np.array_equal(I, Iw)
    assert np.array_equal(D, Dw)


def test_near_ties_are_resolved_exactly(engine, oracle, synth):
    """Vectors that differ from each other by ~1 fp32 ulp of score: the fp16 scan cannot order
    them, the exact band re-rank must."""
    lib = synth.make_library(400, decoy_fraction=0.0, seed=131)
    x = oracle.vectorize(lib["mz"], lib["inten"], lib["off"])
    rng = np.random.default_rng(5)
    base = x[:20].copy()
    clones = []
    for r in range(30):  # 600 near-duplicates of 20 vectors, perturbed in one slot by ~1e-7 relative
        c = base.copy()
        for i in range(20):
            nz = np.flatnonzero(c[i])
            j = nz[rng.integers(len(nz))]
            c[i, j] = np.nextafter(c[i, j], np.float32(2.0 if rng.random() < 0.5 else 0.0))
        clones.append(c)
    x = np.concatenate([x] + clones)
    cent = oracle.kmeans(x, 8, iters=2)
    engine.ivf_set_centroids(CH, cent)
    engine.ivf_add(CH, x)
    off, ids, vecs = oracle.build_lists(x, oracle.ivf_assign(x, cent), 8)
    qv = np.ascontiguousarray(base)
    for k in (5, 17, 31):
        D, I = engine.ivf_search(CH, qv, k, 8)
        Dw, Iw = oracle.ivf_search(qv, cent, off, ids, vecs, 8, k)
        assert np.array_equal(I, Iw) and np.array_equal(D, Dw)


def test_signed_vectors_use_the_norm_bound(engine, oracle):
    """Generic inner-product index with negative entries (not the spectrum use case)."""
    rng = np.random.default_rng(7)
    x = rng.normal(size=(1500, 800)).astype(np.float32)
    x[rng.random(x.shape) < 0.9] = 0
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    q = x[:60] + 0.05 * rng.normal(size=(60, 800)).astype(np.float32)
    cent = oracle.kmeans(x, 8, iters=2)
    engine.ivf_set_centroids(CH, cent)
    engine.ivf_add(CH, x)
    off, ids, vecs = oracle.build_lists(x, oracle.ivf_assign(x, cent), 8)
    D, I = engine.ivf_search(CH, q, 40, 4)
    Dw, Iw = oracle.ivf_search(q, cent, off, ids, vecs, 4, 40)
    assert np.array_equal(I, Iw) and np.array_equal(D, Dw)
