"""Faiss ".idxann" import/export through the C-ABI (reference spectral_library.py:181, :490): files
written by the device index parse with the oracle's NumPy reader; files written by the oracle load
with their lists untouched and search bit-for-bit like the oracle IVF over the same lists."""
import os

import numpy as np
import pytest

from oracle import faiss_io

pytestmark = pytest.mark.gpu
CH = 70


def _world(oracle, synth, n=3000, nlist=24, seed=91):
    lib = synth.make_library(n, decoy_fraction=0.25, seed=seed, decoy_seed=seed + 1)
    x = oracle.vectorize(lib["mz"], lib["inten"], lib["off"])
    cent = oracle.kmeans(x, nlist, seed=4, iters=3)
    q = synth.make_queries(lib, 150, seed=seed + 2)
    qv = oracle.vectorize(q["mz"], q["inten"], q["off"])
    return x, cent, qv


def test_write_index_parses_with_the_oracle_reader(engine, oracle, synth, tmp_path):
    x, cent, qv = _world(oracle, synth)
    x[17] = np.nan  # a row add() skips keeps its id and appears in no list
    engine.ivf_set_centroids(CH, cent)
    engine.ivf_add(CH, x)
    p = str(tmp_path / "w.idxann")
    engine.ivf_write_index(CH, p, nprobe=1)
    r = faiss_io.read_ivf_flat(p)
    assert r["bytes_parsed"] == r["file_size"] == os.path.getsize(p)
    assert (r["fourcc"], r["quantizer_fourcc"], r["d"], r["ntotal"], r["nlist"], r["nprobe"], r["metric"]) == \
           ("IwFl", "IxFI", 800, len(x), 24, 1, 0)
    assert np.array_equal(r["centroids"], cent)
    assign = engine.ivf_assignment(CH)
    assert assign[17] == -1
    for l in range(24):
        rows = np.flatnonzero(assign == l)
        assert np.array_equal(r["list_ids"][l], rows)          # insertion order inside a list
        assert np.array_equal(r["list_vecs"][l], x[rows])      # float32 codes, bit for bit
    from ann_solo_b200.index import inspect_index
    info = inspect_index(p)
    assert info["nstored"] == len(x) - 1 and info["bytes_parsed"] == os.path.getsize(p)


def test_read_index_roundtrip_search_identical(engine, oracle, synth, tmp_path):
    x, cent, qv = _world(oracle, synth)
    engine.ivf_set_centroids(CH, cent)
    engine.ivf_add(CH, x)
    D0, I0 = engine.ivf_search(CH, qv, 64, 6)
    p = str(tmp_path / "r.idxann")
    engine.ivf_write_index(CH, p, nprobe=6)
    assert engine.ivf_read_index(CH + 1, p) == 6
    assert engine.ivf_info(CH + 1) == (len(x), 24, 800)
    assert np.array_equal(engine.ivf_assignment(CH + 1), engine.ivf_assignment(CH))
    assert np.array_equal(engine.ivf_get_centroids(CH + 1), cent)
    assert np.array_equal(engine.ivf_reconstruct(CH + 1), x)
    D1, I1 = engine.ivf_search(CH + 1, qv, 64, 6)
    assert np.array_equal(I0, I1) and np.array_equal(D0, D1)
    # written again from the imported index: the same bytes
    p2 = str(tmp_path / "r2.idxann")
    engine.ivf_write_index(CH + 1, p2, nprobe=6)
    assert open(p, "rb").read() == open(p2, "rb").read()


@pytest.mark.parametrize("sparse", [False, True])
def test_foreign_assignment_is_kept(engine, oracle, synth, tmp_path, sparse):
    """A file whose lists do NOT follow the arg-max-centroid rule (as a Faiss-trained index need not,
    to the last bit): the import keeps the stored lists and the search equals the oracle on them."""
    x, cent, qv = _world(oracle, synth, n=2000, nlist=40, seed=95)
    rng = np.random.default_rng(3)
    lists = np.arange(40) if not sparse else np.arange(0, 40, 3)   # "sprs" encoding: most lists empty
    assign = rng.choice(lists, len(x)).astype(np.int32)
    order = [rng.permutation(np.flatnonzero(assign == l)) for l in range(40)]  # arbitrary order inside a list
    p = str(tmp_path / "f.idxann")
    faiss_io.write_ivf_flat(p, cent, order, [x[i] for i in order], ntotal=len(x) + 5, nprobe=9)
    assert faiss_io.read_ivf_flat(p)["size_encoding"] == ("sprs" if sparse else "full")
    assert engine.ivf_read_index(CH + 2, p) == 9
    got = engine.ivf_assignment(CH + 2)
    assert len(got) == len(x) + 5 and np.array_equal(got[:len(x)], assign) and (got[len(x):] == -1).all()
    off, ids, vecs = oracle.build_lists(x, assign, 40)
    for nprobe, k in [(5, 32), (40, 200)]:
        D, I = engine.ivf_search(CH + 2, qv, k, nprobe)
        Dw, Iw = oracle.ivf_search(qv, cent, off, ids, vecs, nprobe, k)
        assert np.array_equal(I, Iw) and np.array_equal(D, Dw)


def test_add_assigned_and_errors(engine, oracle, synth, tmp_path):
    x, cent, qv = _world(oracle, synth, n=1200, nlist=16, seed=97)
    rng = np.random.default_rng(4)
    assign = rng.integers(-1, 16, len(x)).astype(np.int32)
    engine.ivf_set_centroids(CH + 3, cent)
    engine.ivf_add_assigned(CH + 3, x[:700], assign[:700])
    engine.ivf_add_assigned(CH + 3, x[700:], assign[700:])
    assert np.array_equal(engine.ivf_assignment(CH + 3), assign)
    off, ids, vecs = oracle.build_lists(x, assign, 16)   # drops the -1 rows, keeps global ids
    D, I = engine.ivf_search(CH + 3, qv, 50, 16)
    Dw, Iw = oracle.ivf_search(qv, cent, off, ids, vecs, 16, 50)
    assert np.array_equal(I, Iw) and np.array_equal(D, Dw)
    with pytest.raises(ValueError, match="outside"):
        engine.ivf_add_assigned(CH + 3, x[:2], np.array([0, 16], np.int32))
    # ids that are not sequential / stored twice are refused
    p = str(tmp_path / "dup.idxann")
    faiss_io.write_ivf_flat(p, cent[:2], [np.array([0, 1]), np.array([1])], [x[:2], x[1:2]])
    with pytest.raises(ValueError, match="stored twice"):
        engine.ivf_read_index(CH + 4, p)
    faiss_io.write_ivf_flat(p, cent[:2], [np.array([0, 7]), np.array([1])], [x[:2], x[1:2]])
    with pytest.raises(ValueError, match="sequential"):
        engine.ivf_read_index(CH + 4, p)
    faiss_io.write_ivf_flat(p, cent[:2], [np.array([0]), np.array([1])], [x[:1], x[1:2]], metric=1)
    with pytest.raises(ValueError, match="METRIC_INNER_PRODUCT"):
        engine.ivf_read_index(CH + 4, p)


def test_faiss_like_module_functions(engine, oracle, synth, tmp_path):
    from ann_solo_b200 import index as faiss
    x, cent, qv = _world(oracle, synth, n=1500, nlist=16, seed=99)
    ix = faiss.IndexIVFFlat(faiss.IndexFlatIP(800), 800, 16, faiss.METRIC_INNER_PRODUCT, engine=engine)
    ix.set_centroids(cent)
    ix.add(x)
    ix.nprobe = 4
    D0, I0 = ix.search(qv, 30)
    p = str(tmp_path / "m.idxann")
    faiss.write_index(ix, p)
    ix2 = faiss.read_index(p, engine=engine)
    assert (ix2.ntotal, ix2.nlist, ix2.d, ix2.nprobe) == (len(x), 16, 800, 4)
    D1, I1 = ix2.search(qv, 30)
    assert np.array_equal(I0, I1) and np.array_equal(D0, D1)
    assert np.array_equal(ix2.reconstruct_n(5, 10), x[5:15])


def test_spectral_library_caches_indexes_in_idxann_files(engine, synth, tmp_path):
    """reference spectral_library.py:92-115: {basename}_{hash[:7]}_{charge}.idxann is written on the
    first construction and read (not rebuilt) on the second; both give the same SSMs."""
    import glob
    from ann_solo_b200.config import config
    from ann_solo_b200.spectral_library import InMemoryLibrary, SpectralLibrary
    lib = synth.make_library(2500, seed=111, decoy_seed=112)
    queries = synth.make_queries(lib, 100, seed=113)
    reader, qreader = InMemoryLibrary(lib), InMemoryLibrary(queries)
    config.update(dict(num_list=16, num_probe=8, num_candidates=64, precursor_tolerance_mass_open=300.0,
                       precursor_tolerance_mode_open="Da"))
    try:
        base = str(tmp_path / "lib")
        sl = SpectralLibrary(reader, engine=engine, train_iters=3, ann_basename=base)
        files = sorted(glob.glob(base + "_*.idxann"))
        assert len(files) == len(sl._ann_charges) >= 2
        assert all(os.path.basename(f).split("_")[1] == sl._get_hyperparameter_hash()[:7] for f in files)
        qs = [qreader.read_spectrum(i) for i in range(100)]
        for s in qs:
            s.is_processed = True
        q2 = [s for s in qs if s.precursor_charge == 2]
        first = [(s.query_identifier, s.library_identifier, s.search_engine_score) for s in sl._search_batch(q2, 2, "open")]
        stamp = [os.path.getmtime(f) for f in files]
        cent = engine.ivf_get_centroids(2)
        sl2 = SpectralLibrary(reader, engine=engine, train_iters=1, ann_basename=base)  # would train differently
        assert [os.path.getmtime(f) for f in files] == stamp
        assert np.array_equal(engine.ivf_get_centroids(2), cent)
        again = [(s.query_identifier, s.library_identifier, s.search_engine_score) for s in sl2._search_batch(q2, 2, "open")]
        assert first == again and len(first) > 10
    finally:
        config.update(dict(num_list=256, num_probe=128, num_candidates=1024))
