"""The IVF restatement (parity unpinned against Faiss — the arithmetic is defined by the oracle):
internal consistency, ordering, tie-break, padding, NaN handling."""
import numpy as np


def _vecs(oracle, synth, n=1500, seed=1):
    lib = synth.make_library(n, decoy_fraction=0.0, seed=seed)
    return oracle.vectorize(lib["mz"], lib["inten"], lib["off"])


def test_ip_is_sequential_fmaf(oracle):
    rng = np.random.default_rng(0)
    a = rng.random(800).astype(np.float32)
    b = rng.random(800).astype(np.float32)
    assert abs(oracle.ip(a, b) - float(np.dot(a.astype(np.float64), b.astype(np.float64)))) < 1e-4
    # zeros can be skipped without changing the result (what the sparse GPU kernels rely on)
    b[rng.random(800) < 0.9] = 0
    nz = np.flatnonzero(b)
    assert oracle.ip(a, b) == oracle.ip(a[nz], b[nz])


def test_search_full_probe_equals_brute_force(oracle, synth):
    x = _vecs(oracle, synth)
    q = x[:40] * 0.5 + _vecs(oracle, synth, 40, seed=9) * 0.5
    cent = oracle.kmeans(x, 16, seed=4, iters=3)
    assign = oracle.ivf_assign(x, cent)
    off, ids, vecs = oracle.build_lists(x, assign, 16)
    D, I = oracle.ivf_search(q, cent, off, ids, vecs, nprobe=16, k=50)
    one = np.zeros((1, 800), np.float32)
    D1, I1 = oracle.ivf_search(q, one, np.array([0, len(x)]), np.arange(len(x)), x, nprobe=1, k=50)
    assert np.array_equal(I, I1) and np.array_equal(D, D1)
    assert (np.diff(D, axis=1) <= 0).all()
    for i in range(5):
        assert D[i, 0] == oracle.ip(q[i], x[I[i, 0]])


def test_assign_matches_coarse_top1_and_ties(oracle, synth):
    x = _vecs(oracle, synth, 600)
    cent = oracle.kmeans(x, 8, iters=2)
    cent[5] = cent[2]  # exact duplicate centroid: ties resolve to the lower id
    assign = oracle.ivf_assign(x, cent)
    probes, _ = oracle.ivf_coarse(x, cent, 1)
    assert np.array_equal(assign, probes[:, 0])
    assert np.array_equal(assign, oracle.ivf_assign(x, cent, fast=False))  # fast variant is bit-identical
    assert not (assign == 5).any()


def test_padding_and_nan_rows(oracle, synth):
    x = _vecs(oracle, synth, 300)
    x[7] = np.nan  # an invalid library spectrum vectorises to 0/0
    cent = oracle.kmeans(x, 4, iters=2)
    assign = oracle.ivf_assign(x, cent)
    assert assign[7] == -1
    off, ids, vecs = oracle.build_lists(x, assign, 4)
    assert off[-1] == 299 and 7 not in ids
    D, I = oracle.ivf_search(x[:3], cent, off, ids, vecs, nprobe=1, k=299)
    assert (I[:, -1] == -1).all() and np.isneginf(D[:, -1]).all()
    for i in range(3):
        valid = I[i][I[i] >= 0]
        assert len(np.unique(valid)) == len(valid)


def test_duplicate_vectors_tie_break_by_id(oracle, synth):
    x = _vecs(oracle, synth, 200)
    x[150] = x[20]
    cent = oracle.kmeans(x, 2, iters=1)
    off, ids, vecs = oracle.build_lists(x, oracle.ivf_assign(x, cent), 2)
    D, I = oracle.ivf_search(x[20:21], cent, off, ids, vecs, nprobe=2, k=2)
    assert I[0].tolist() == [20, 150] and D[0, 0] == D[0, 1]


def test_simd_variant_close_to_exact(oracle, synth):
    x = _vecs(oracle, synth, 1200)
    cent = oracle.kmeans(x, 16, iters=2)
    off, ids, vecs = oracle.build_lists(x, oracle.ivf_assign(x, cent), 16)
    D, I = oracle.ivf_search(x[:30], cent, off, ids, vecs, 4, 20)
    D2, I2 = oracle.ivf_search(x[:30], cent, off, ids, vecs, 4, 20, simd=True)
    np.testing.assert_allclose(D, D2, rtol=1e-5, atol=1e-7)
    assert (I == I2).mean() > 0.95


def test_candidates_post_filter_semantics(oracle):
    # window applied AFTER the top-k (SURVEY.md finding 5)
    lib_mz = np.array([500.0, 500.2, 600.0, 700.0], np.float32)
    valid = np.array([1, 1, 1, 0], np.uint8)
    ann = np.array([[2, 0, -1]], np.int64)
    ids, off = oracle.candidates([500.1], lib_mz, valid, 2, 1.0, "Da", ann)
    assert ids.tolist() == [0] and off.tolist() == [0, 1]
    ids, off = oracle.candidates([500.1], lib_mz, valid, 2, 1.0, "Da", None)
    assert ids.tolist() == [0, 1]
    ids, off = oracle.candidates([500.0001], lib_mz, valid, 2, 20.0, "ppm", None)
    assert ids.tolist() == [0]
