"""The fused open-search batch (K1 -> K2/K3/K4 -> window mask -> K5) vs the same pipeline
assembled from oracle calls; brute-force mode; the SpectralLibrary drop-in surface."""
import numpy as np
import pytest

from conftest import canon_pairs

pytestmark = pytest.mark.gpu


def _oracle_pipeline(oracle, store, q, cent, nlist, charge, k, nprobe, tol, tol_mode, use_ann=True):
    x = oracle.vectorize(store["mz"], store["inten"], store["off"])
    ann = None
    if use_ann:
        assign = oracle.ivf_assign(x, cent)
        off, ids, vecs = oracle.build_lists(x, assign, nlist)
        qv = oracle.vectorize(q["mz"], q["inten"], q["off"])
        _, ann = oracle.ivf_search(qv, cent, off, ids, vecs, nprobe, k)
    cand, coff = oracle.candidates(q["prec_mz"], store["prec_mz"].astype(np.float32), store["valid"], charge, tol,
                                   tol_mode, ann)
    bp, bs, npairs, pairs = oracle.best_match_batch(q, store, cand, coff, 0.02, True, sort_mode=1)
    row = np.full(len(bp), -1, np.int32)
    has = bp >= 0
    row[has] = cand[coff[:-1][has] + bp[has]]
    return row, bs, npairs, pairs, np.diff(coff)


@pytest.mark.parametrize("use_ann,tol,tol_mode", [(True, 500.0, "Da"), (True, 20.0, "ppm"), (False, 5.0, "Da"),
                                                 (False, 20.0, "ppm")])
def test_fused_batch_equals_oracle_pipeline(engine, oracle, synth, small_world, use_ann, tol, tol_mode):
    from ann_solo_b200.engine import SoloEngine
    lib, per_charge, queries = small_world
    for charge in (2, 3):
        store, rows = per_charge[charge]
        store = dict(store)
        store["valid"] = store["valid"].copy()
        store["valid"][::17] = 0  # some invalid library spectra (dropped after the top-k, :453)
        qsel = np.flatnonzero(queries["prec_z"] == charge)
        q = synth.take_spectra(queries, qsel)
        nlist = 16
        x = oracle.vectorize(store["mz"], store["inten"], store["off"])
        cent = oracle.kmeans(x, nlist, iters=3)
        engine.set_vectorizer(11, 2010, 0.04, 800)
        engine.load_library(charge, store)
        engine.ivf_set_centroids(charge, cent)
        engine.ivf_add_library(charge)
        assert np.array_equal(engine.ivf_assignment(charge), oracle.ivf_assign(x, cent))
        p = SoloEngine.make_params(use_ann, 64, 6, tol, tol_mode, 0.02, True, max_pairs=50)
        res = engine.search_batch(charge, p, q)
        row, bs, npairs, pairs, ncand = _oracle_pipeline(oracle, store, q, cent, nlist, charge, 64, 6, tol, tol_mode,
                                                         use_ann)
        assert np.array_equal(res["n_cand"], ncand)
        assert np.array_equal(res["best_row"], row)
        has = row >= 0
        assert has.sum() > 5
        assert np.array_equal(res["score"][has], bs[has])
        assert np.array_equal(res["n_pairs"][has], npairs[has])
        for i in np.flatnonzero(has):
            n = npairs[i]
            assert np.array_equal(res["pairs"][i, :n], pairs[i, :n])


def test_f64_query_mz_binning(engine, oracle, synth, small_world):
    """Queries held as float64 m/z are binned in float64 (NumPy semantics) but scored as float32."""
    from ann_solo_b200.engine import SoloEngine
    lib, per_charge, queries = small_world
    store, _ = per_charge[2]
    q = synth.take_spectra(queries, np.flatnonzero(queries["prec_z"] == 2)[:60])
    mz64 = q["mz"].astype(np.float64) + 3e-6
    x = oracle.vectorize(store["mz"], store["inten"], store["off"])
    cent = oracle.kmeans(x, 16, iters=3)
    engine.load_library(2, store)
    engine.ivf_set_centroids(2, cent)
    engine.ivf_add_library(2)
    p = SoloEngine.make_params(True, 32, 4, 500.0, "Da", 0.02, True, max_pairs=50, mz_is_f64=True)
    res = engine.search_batch(2, p, q, mz_vec=mz64)
    off, ids, vecs = oracle.build_lists(x, oracle.ivf_assign(x, cent), 16)
    qv = oracle.vectorize(mz64, q["inten"], q["off"])
    _, ann = oracle.ivf_search(qv, cent, off, ids, vecs, 4, 32)
    cand, coff = oracle.candidates(q["prec_mz"], store["prec_mz"].astype(np.float32), store["valid"], 2, 500.0, "Da", ann)
    bp, bs, _, _ = oracle.best_match_batch(q, store, cand, coff, 0.02, True, sort_mode=1)
    assert np.array_equal(res["n_cand"], np.diff(coff))
    has = bp >= 0
    assert np.array_equal(res["best_row"][has], cand[coff[:-1][has] + bp[has]])
    assert np.array_equal(res["score"][has], bs[has])


def test_spectral_library_dropin(engine, oracle, synth):
    """SpectralLibrary.search / _search_batch / _get_library_candidates keep the reference's
    surface; the fused batch and the explicit candidate lists agree with each other and with
    the oracle-driven flow."""
    from ann_solo_b200.config import config
    from ann_solo_b200.spectral_library import InMemoryLibrary, SpectralLibrary
    from ann_solo_b200.spectrum_match import get_best_match
    lib = synth.make_library(2500, seed=101, decoy_seed=102)
    queries = synth.make_queries(lib, 120, seed=103)
    reader, qreader = InMemoryLibrary(lib), InMemoryLibrary(queries)
    config.update(dict(num_list=16, num_probe=8, num_candidates=64, precursor_tolerance_mass=20.0,
                       precursor_tolerance_mode="ppm", precursor_tolerance_mass_open=300.0,
                       precursor_tolerance_mode_open="Da", fragment_mz_tolerance=0.02, allow_peak_shifts=True,
                       fdr=0.05))
    try:
        sl = SpectralLibrary(reader, engine=engine, train_iters=3)
        qs = [qreader.read_spectrum(i) for i in range(120)]
        for s in qs:
            s.is_processed = True
        q2 = [s for s in qs if s.precursor_charge == 2]
        # fused batch vs explicit candidates + per-query get_best_match
        ssms = {s.query_identifier: s for s in sl._search_batch(q2, 2, "open")}
        cands = list(sl._get_library_candidates(q2, 2, "open"))
        assert len(cands) == len(q2)
        n_checked = 0
        for query, cl in zip(q2, cands):
            if not cl:
                assert query.identifier not in ssms
                continue
            match, score, pm = get_best_match(query, cl, 0.02, True, engine=engine)
            ssm = ssms[query.identifier]
            assert ssm.library_identifier == match.identifier
            assert ssm.search_engine_score == score
            assert np.array_equal(canon_pairs(ssm.peak_matches, len(pm)), canon_pairs(np.array(pm), len(pm)))
            assert ssm.peak_matches.ndim == 2 and ssm.peak_matches.shape[1] == 2
            n_checked += 1
        assert n_checked > 20
        engine.load_library(2, reader.charge_store(2))  # get_best_match used the scratch slot only
        with pytest.raises(ValueError):
            list(sl._search_batch(q2, 2, "bogus"))
        # whole cascade
        out = sl.search(qs)
        ident = {s.query_identifier: s for s in out}
        truth = queries["truth"]
        correct = sum(1 for i, s in ident.items() if truth[i] >= 0 and s.library_identifier == truth[i])
        assert len(out) > 60 and correct > 40
        sl.shutdown()
    finally:
        config.update(dict(num_list=256, num_probe=128, num_candidates=1024))


@pytest.mark.parametrize("tol,tol_mode", [(0.5, "Da"), (20.0, "ppm"), (0.0, "Da"), (2.0e6, "ppm")])
def test_window_only_candidates_at_the_boundaries(engine, oracle, synth, tol, tol_mode):
    """Brute-force candidate sets (level-1 'std' search, --mode bf) from the m/z-sorted view: precursors
    exactly on the window edge, duplicates, invalid rows and a NaN query must give the oracle's counts."""
    from ann_solo_b200.engine import SoloEngine
    charge = 2
    lib = synth.make_library(1500, seed=71, decoy_seed=72)
    store, _ = synth.split_by_charge(lib)[charge]
    store = dict(store)
    n = len(store["prec_mz"])
    rng = np.random.default_rng(5)
    store["prec_mz"] = store["prec_mz"].copy()
    store["prec_mz"][:50] = 700.0                      # duplicates
    store["prec_mz"][50:60] = np.float64(np.float32(700.25))   # exactly tol / charge away for the Da case
    store["valid"] = store["valid"].copy()
    store["valid"][rng.integers(0, n, 100)] = 0
    queries = synth.make_queries(lib, 200, seed=73)
    q = synth.take_spectra(queries, np.flatnonzero(queries["prec_z"] == charge))
    q["prec_mz"] = q["prec_mz"].copy()
    q["prec_mz"][0] = 700.0
    q["prec_mz"][1] = 700.0 + 0.25
    q["prec_mz"][2] = 700.0 * (1 + 20e-6)
    q["prec_mz"][3] = np.nan
    q["prec_mz"][4] = 1.0e9
    engine.load_library(charge, store)
    p = SoloEngine.make_params(False, 64, 6, tol, tol_mode, 0.02, True, max_pairs=50)
    res = engine.search_batch(charge, p, q)
    cand, coff = oracle.candidates(q["prec_mz"], store["prec_mz"].astype(np.float32), store["valid"], charge, tol, tol_mode)
    assert np.array_equal(res["n_cand"], np.diff(coff))
    bp, bs, npairs, pairs = oracle.best_match_batch(q, store, cand, coff, 0.02, True, sort_mode=1)
    has = bp >= 0
    row = np.full(len(bp), -1, np.int32)
    row[has] = cand[coff[:-1][has] + bp[has]]
    assert np.array_equal(res["best_row"], row)
    assert np.array_equal(res["score"][has], bs[has])
