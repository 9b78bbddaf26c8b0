"""K1 (CUDA) vs the oracle and the reference's golden vectors. Bit-exact against the oracle;
<= 3e-7 relative against NumPy's unspecified BLAS norm order in the reference fixtures."""
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def test_device_hash_lut_matches_reference(engine):
    g = np.load(os.path.join(GOLDEN, "vectoriser.npz"))
    engine.set_vectorizer(11, 2010, 0.04, 800)
    for b in list(range(0, 49978, 997)) + [0, 1, 2225, 2226, 49976, 49977]:
        assert engine.hash_slot(b) == int(g["hash_800"][b])


@pytest.mark.parametrize("prec", ["f32", "f64"])
def test_golden_vectors(engine, oracle, prec):
    g = np.load(os.path.join(GOLDEN, "vectoriser.npz"))
    engine.set_vectorizer(11, 2010, 0.04, 800)
    mz = g["mz64"].astype(np.float32) if prec == "f32" else g["mz64"]
    got = engine.vectorize(mz, g["inten"], g["off"])
    want = g["v32"] if prec == "f32" else g["v64"]
    assert np.array_equal(got != 0, want != 0)
    np.testing.assert_allclose(got, want, rtol=3e-7, atol=0)
    assert np.array_equal(got, oracle.vectorize(mz, g["inten"], g["off"]))  # bit-exact vs oracle
    if prec == "f64":
        assert np.array_equal(engine.vectorize(mz, g["inten"], g["off"], norm=False), g["v64_raw"])


@pytest.mark.parametrize("hash_len,bin_size", [(800, 0.04), (400, 0.05), (1024, 1.0005)])
def test_random_batches_bit_exact(engine, oracle, synth, hash_len, bin_size):
    lib = synth.make_library(3000, seed=31, decoy_seed=32)
    engine.set_vectorizer(11, 2010, bin_size, hash_len)
    try:
        for mz in (lib["mz"], lib["mz"].astype(np.float64) + 1e-7):
            got = engine.vectorize(mz, lib["inten"], lib["off"])
            want = oracle.vectorize(mz, lib["inten"], lib["off"], 11, 2010, bin_size, hash_len)
            assert np.array_equal(got, want)
            assert np.allclose(np.linalg.norm(got.astype(np.float64), axis=1), 1.0, atol=1e-6)
    finally:
        engine.set_vectorizer(11, 2010, 0.04, 800)


def test_edge_cases(engine, oracle):
    engine.set_vectorizer(11, 2010, 0.04, 800)
    # empty spectrum -> 0/0 = NaN row like NumPy; ragged sizes; out-of-range m/z (hashed on device)
    mz = np.array([100.0, 5.0, 2500.0, 300.0, 300.01, 300.02], np.float32)
    inten = np.array([1.0, 2.0, 3.0, 0.5, 0.25, 0.125], np.float32)
    off = np.array([0, 0, 1, 3, 6, 6], np.int64)
    got = engine.vectorize(mz, inten, off)
    want = oracle.vectorize(mz, inten, off)
    assert np.isnan(got[0]).all() and np.isnan(got[4]).all()
    assert np.array_equal(np.isnan(got), np.isnan(want))
    assert np.array_equal(np.nan_to_num(got), np.nan_to_num(want))
    # many peaks in one spectrum (collisions add in peak order)
    rng = np.random.default_rng(1)
    mz = np.sort(rng.uniform(11, 2010, 5000)).astype(np.float32)
    inten = rng.random(5000).astype(np.float32)
    off = np.array([0, 5000], np.int64)
    assert np.array_equal(engine.vectorize(mz, inten, off), oracle.vectorize(mz, inten, off))
    assert engine.vectorize(mz[:0], inten[:0], np.array([0], np.int64)).shape == (0, 800)


def test_spectrum_to_vector_api(engine, oracle):
    from ann_solo_b200.spectrum import MsmsSpectrum, spectrum_to_vector
    g = np.load(os.path.join(GOLDEN, "vectoriser.npz"))
    b, e = g["off"][3], g["off"][4]
    s = MsmsSpectrum("s", 500.0, 2, g["mz64"][b:e], g["inten"][b:e])
    v = spectrum_to_vector(s, 11, 2010, 0.04, 800, engine=engine)
    np.testing.assert_allclose(v, g["v64"][3], rtol=3e-7)
    buf = np.zeros(800, np.float32)
    assert spectrum_to_vector(s, 11, 2010, 0.04, 800, True, buf, engine=engine) is buf
    with pytest.raises(ValueError):
        spectrum_to_vector(s, 11, 2010, 0.04, 800, True, np.zeros(10, np.float32), engine=engine)
