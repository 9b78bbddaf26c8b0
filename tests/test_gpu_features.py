"""K6 (SSM feature table) through the C-ABI vs the golden rows minted from the reference's
spectrum_similarity.py and vs the oracle restatement; tolerance as in tests/test_ssm_features.py."""
import numpy as np
import pytest

from oracle import ssm_features as sf
from test_ssm_features import BINS, NAMES, assert_rows_close, golden_cases

pytestmark = pytest.mark.gpu
CH = 90


def _stores(cases):
    def csr(key_mz, key_int, prec):
        off = np.cumsum([0] + [len(c[key_int]) for c in cases]).astype(np.int64)
        return dict(mz=np.concatenate([np.asarray(c[key_mz], np.float32) for c in cases]),
                    mz64=np.concatenate([np.asarray(c[key_mz], np.float64) for c in cases]),
                    inten=np.concatenate([c[key_int] for c in cases]).astype(np.float32), off=off,
                    prec_mz=np.array([c[prec] for c in cases], np.float64),
                    prec_z=np.full(len(cases), 2, np.int32), chg=None, valid=None)
    return csr("q_mz", "q_int", "q_prec"), csr("l_mz", "l_int", "l_prec")


def _pairs(cases):
    mp = max(len(c["pairs"]) for c in cases)
    pairs = np.zeros((len(cases), mp, 2), np.uint32)
    for i, c in enumerate(cases):
        pairs[i, :len(c["pairs"])] = c["pairs"]
    return pairs, np.array([len(c["pairs"]) for c in cases], np.int32)


@pytest.mark.parametrize("precision", ["f32", "f64"])
def test_golden_rows_through_the_cabi(engine, precision):
    cases = [c for c in golden_cases() if (c["q_mz"].dtype == np.float32) == (precision == "f32")]
    assert len(cases) > 30
    q, lib = _stores(cases)
    engine.set_vectorizer(11, 2010, 0.04, 800)
    engine.load_library(CH, lib)
    pairs, n_pairs = _pairs(cases)
    q_charge = np.array([c["q_z"] for c in cases], np.int32)
    seq_len = np.arange(len(cases), dtype=np.int32) % 30
    got = engine.ssm_features(CH, q, np.arange(len(cases)), pairs, n_pairs, q_charge, seq_len,
                              mz_vec=q["mz64"] if precision == "f64" else None)
    assert got.shape == (len(cases), 44)
    assert np.array_equal(got[:, 0], seq_len)
    got[:, 0] = 0
    for i, c in enumerate(cases):
        assert_rows_close(got[i], c["row"], c["name"])
        want = sf.ssm_features(c["q_mz"], c["q_int"], c["l_mz"], c["l_int"], c["pairs"], c["q_prec"], c["q_z"],
                               c["l_prec"], 0, BINS)
        np.testing.assert_allclose(got[i], want, rtol=1e-9, atol=1e-11, err_msg=c["name"])


def test_skipped_rows_and_argument_checks(engine):
    cases = list(golden_cases())[:6]
    q, lib = _stores(cases)
    engine.load_library(CH, lib)
    pairs, n_pairs = _pairs(cases)
    rows = np.arange(6, dtype=np.int32)
    rows[2] = -1            # no library match
    n_pairs[4] = 0          # no peak matches: skipped by the reference (utils.py:332-333)
    got = engine.ssm_features(CH, q, rows, pairs, n_pairs)
    assert np.isnan(got[2]).all() and np.isnan(got[4]).all() and np.isfinite(got[[0, 1, 3, 5], :20]).all()
    assert (got[[0, 1, 3, 5], 1:5].sum(axis=1) == 1).all() and (got[[0, 1, 3, 5], 4] == 1).all()  # charge 90 -> ">= 5"
    with pytest.raises(ValueError, match="library row"):
        engine.ssm_features(CH, q, np.full(6, 99, np.int32), pairs, n_pairs)
    bad = pairs.copy()
    bad[0, 0, 0] = 200
    with pytest.raises(ValueError, match="query peak"):
        engine.ssm_features(CH, q, rows, bad, n_pairs)
    from ann_solo_b200 import SoloError
    bad = pairs.copy()
    bad[1, 0, 1] = 120          # library peak index beyond the matched library spectrum: caught on the device
    with pytest.raises(SoloError, match="outside its spectra"):
        engine.ssm_features(CH, q, rows, bad, n_pairs)


def test_staged_features_equal_oracle_on_a_fused_search(engine, oracle, synth, small_world):
    lib, per_charge, queries = small_world
    L, _ = per_charge[2]
    engine.set_vectorizer(11, 2010, 0.04, 800)
    engine.load_library(2, L)
    qsel = np.flatnonzero(queries["prec_z"] == 2)
    q = synth.take_spectra(queries, qsel)
    params = engine.make_params(False, 64, 8, 300.0, "Da", 0.02, True, max_pairs=64)
    res = engine.search_batch(2, params, q)
    seq_len = (np.arange(len(qsel)) % 17 + 6).astype(np.int32)
    got = engine.ssm_features_staged(2, None, seq_len)
    want = sf.ssm_features_batch(q, L, res["best_row"], res["pairs"], res["n_pairs"], np.full(len(qsel), 2), seq_len,
                                 BINS)
    has = (res["best_row"] >= 0) & (res["n_pairs"] > 0)
    assert has.sum() > 50
    assert np.isnan(got[~has]).all()
    np.testing.assert_allclose(got[has], want[has], rtol=1e-9, atol=1e-11)
    # the host-buffer entry point gives the same table
    again = engine.ssm_features(2, q, res["best_row"], res["pairs"], res["n_pairs"], None, seq_len)
    assert np.array_equal(got[has], again[has])
    # properties that hold for every SSM
    f = dict(zip(NAMES, got[has].T))
    assert (f["n_matched_peaks"] == res["n_pairs"][has]).all()
    assert ((f["cosine"] > 0) & (f["cosine"] <= 1 + 1e-6)).all() and (f["precursor_charge_2"] == 1).all()
    # the search score down-weights shifted matches (SpectrumMatch.cpp), the cosine feature does not
    assert (f["cosine"] >= res["score"][has] * (1 - 1e-6)).all()


def test_compute_ssm_features_mirror(engine, synth):
    """utils._compute_ssm_features keeps the reference's dictionary (utils.py:296-340, :455-457)."""
    from ann_solo_b200.config import config
    from ann_solo_b200.spectral_library import InMemoryLibrary, SpectralLibrary
    lib = synth.make_library(2000, seed=121, decoy_seed=122)
    queries = synth.make_queries(lib, 90, seed=123)
    peptides = ["PEPTIDEK" + "A" * (i % 9) for i in range(len(lib["prec_mz"]))]
    reader, qreader = InMemoryLibrary(lib, peptides=peptides), InMemoryLibrary(queries)
    config.update(dict(mode="bf", precursor_tolerance_mass_open=300.0, precursor_tolerance_mode_open="Da"))
    try:
        sl = SpectralLibrary(reader, engine=engine)
        qs = [qreader.read_spectrum(i) for i in range(90)]
        for s in qs:
            s.is_processed = True
        ssms = []
        for z in (2, 3):
            ssms += list(sl._search_batch([s for s in qs if s.precursor_charge == z], z, "open"))
        ssms.append(type(ssms[0])(ssms[0].query_spectrum, ssms[0].library_spectrum, np.zeros((0, 2), np.int64)))
        feats = sl.compute_ssm_features(ssms)
        assert list(feats)[:2] == ["index", "sequence"] and list(feats)[-1] == "is_target" and len(feats) == 47
        assert list(feats)[2:-1] == NAMES
        n = len(feats["index"])
        with_matches = [i for i, s in enumerate(ssms) if len(s.peak_matches) > 0]   # others are skipped (:332-333)
        assert feats["index"] == with_matches and len(ssms) - 1 not in with_matches and n > 60
        assert all(len(v) == n for v in feats.values())
        for j, i in enumerate(feats["index"]):
            ssm = ssms[i]
            want = sf.ssm_features(ssm.query_spectrum.mz, ssm.query_spectrum.intensity, ssm.library_spectrum.mz,
                                   ssm.library_spectrum.intensity, ssm.peak_matches, ssm.query_spectrum.precursor_mz,
                                   ssm.query_spectrum.precursor_charge, ssm.library_spectrum.precursor_mz,
                                   len(ssm.sequence), BINS)
            got = np.array([feats[k][j] for k in NAMES], np.float64)
            np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-11)
            assert feats["sequence"][j] == ssm.sequence and feats["is_target"][j] == (not ssm.is_decoy)
            assert isinstance(feats["n_matched_peaks"][j], int) and feats["sequence_len"][j] == len(ssm.sequence)
    finally:
        config.update(dict(mode="ann"))
