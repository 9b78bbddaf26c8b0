"""K0 (batched process_spectrum) through the C-ABI vs the oracle restatement process_spectrum_np:
kept peaks, m/z, float32 intensities and validity bit for bit."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _raw(rng, n, f64=False, ties=False):
    counts = rng.integers(0, 400, n)
    counts[:4] = [0, 5, 9, 3000]
    off = np.zeros(n + 1, np.int64)
    np.cumsum(counts, out=off[1:])
    mz = np.concatenate([np.sort(rng.uniform(5.0, 2100.0 if i % 3 else 240.0, c)) for i, c in enumerate(counts)])
    inten = rng.gamma(0.6, 1000.0, off[-1]).astype(np.float32)
    if ties:
        inten = np.round(inten / 200).astype(np.float32) * 200 + 1
    prec_z = rng.integers(1, 5, n).astype(np.int32)
    prec_mz = rng.uniform(300, 1000, n)
    for i in range(0, n, 5):   # plant peaks at the precursor and its isotopes / lower charge states
        b, e = off[i], off[i + 1]
        if e - b > 30:
            neutral = (prec_mz[i] - 1.0072766) * prec_z[i]
            for k, (c, iso) in enumerate([(prec_z[i], 0), (1, 1), (max(prec_z[i] - 1, 1), 2)]):
                mz[b + 10 + k] = (neutral + iso) / c + 1.0072766 + rng.uniform(-0.04, 0.04)
            mz[b:e] = np.sort(mz[b:e])
    store = dict(mz=mz.astype(np.float32), inten=inten, off=off, prec_mz=prec_mz, prec_z=prec_z,
                 chg=rng.integers(0, 3, off[-1]).astype(np.uint8))
    return store, (mz if f64 else None)


CONFIGS = [
    dict(),
    dict(scaling="root"),
    dict(scaling=None, max_peaks=30),
    dict(remove_precursor=True, remove_precursor_tolerance=0.05),
    dict(remove_precursor=True, remove_precursor_tolerance=1.5, scaling="sqrt", min_intensity=0.05, max_peaks=128),
    dict(min_mz=101.5, max_mz=1500.25, min_peaks=20, min_mz_range=400.5, max_peaks=50),
    # round(resolution, 'sum') (reference spectrum.py:84-89): whole-number m/z merges many peaks, 1 and 2 decimals few
    dict(resolution=0),
    dict(resolution=1, remove_precursor=True, remove_precursor_tolerance=0.5, scaling="root"),
    dict(resolution=2, scaling=None, max_peaks=64, min_intensity=0.02),
]


@pytest.mark.parametrize("cfg", CONFIGS)
@pytest.mark.parametrize("f64,ties", [(False, False), (True, False), (False, True)])
def test_process_spectra_bit_exact(engine, oracle, cfg, f64, ties):
    rng = np.random.default_rng(17 + len(cfg))
    store, mz64 = _raw(rng, 160, f64, ties)
    got = engine.process_spectra(store, mz_vec=mz64, **cfg)
    mz_in = mz64 if f64 else store["mz"]
    off = store["off"]
    n_valid = 0
    for i in range(160):
        b, e = off[i], off[i + 1]
        w_mz, w_int, w_valid, w_idx = oracle.process_spectrum_np(mz_in[b:e], store["inten"][b:e],
                                                                 float(store["prec_mz"][i]), int(store["prec_z"][i]),
                                                                 **cfg)
        assert bool(got["valid"][i]) == w_valid, i
        gb, ge = got["off"][i], got["off"][i + 1]
        if not w_valid:
            assert ge == gb
            continue
        n_valid += 1
        assert np.array_equal(got["src"][gb:ge], w_idx), i
        assert np.array_equal((got["mz64"] if f64 else got["mz"])[gb:ge], w_mz), i
        assert np.array_equal(got["inten"][gb:ge].view(np.uint32), w_int.view(np.uint32)), i
        assert np.array_equal(got["chg"][gb:ge], store["chg"][b:e][w_idx]), i
        assert ge - gb <= cfg.get("max_peaks", 50)
    assert 40 < n_valid < 160


def test_processed_spectra_are_fixed_points_and_errors(engine, oracle):
    rng = np.random.default_rng(5)
    store, _ = _raw(rng, 64)
    once = engine.process_spectra(store, scaling="rank")
    assert np.allclose(np.add.reduceat(once["inten"].astype(np.float64) ** 2, once["off"][:-1][once["valid"] > 0]), 1,
                       atol=1e-6)
    # unit norm, at most 50 peaks, m/z ascending inside [11, 2010]
    for i in np.flatnonzero(once["valid"]):
        m = once["mz"][once["off"][i]:once["off"][i + 1]]
        assert 10 <= len(m) <= 50 and (np.diff(m) >= 0).all() and m[0] >= 11 and m[-1] <= 2010
    with pytest.raises(ValueError, match="resolution"):
        engine.process_spectra(store, resolution=13)
    with pytest.raises(ValueError, match="scaling"):
        engine.process_spectra(store, scaling="log")
    from ann_solo_b200 import SoloError
    with pytest.raises(SoloError, match="max_peaks"):
        engine.process_spectra(store, max_peaks=200)
    big = dict(mz=np.linspace(100, 1900, 9000).astype(np.float32), inten=np.ones(9000, np.float32),
               off=np.array([0, 9000]), prec_mz=np.array([500.0]), prec_z=np.array([2], np.int32))
    with pytest.raises(SoloError, match="8192"):
        engine.process_spectra(big)
