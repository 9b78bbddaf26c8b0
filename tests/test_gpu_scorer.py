"""K5 (CUDA) vs the oracle (total-order sort: bit-exact including the order of the pairs), the
reference's compiled core (pairs compared canonically) and the golden fixtures."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, canon_pairs
from test_oracle_golden import _random_fixture, _store

pytestmark = pytest.mark.gpu
CH = 77  # scratch slot of the engine's library table


def test_kats(engine):
    kats = json.load(open(os.path.join(GOLDEN, "scorer_kat.json")))
    for k in kats:
        q, lib = _store([k["query"]]), _store(k["candidates"])
        engine.load_library(CH, lib)
        n = len(k["candidates"])
        bp, bs, npairs, pairs = engine.best_match_batch(CH, q, np.arange(n), np.array([0, n]), k["tol"],
                                                        k["allow_shift"])
        assert bp[0] == k["best"], k["name"]
        assert bs[0] == k["score"], k["name"]
        assert pairs[0, :npairs[0]].tolist() == k["pairs"], k["name"]


@pytest.mark.parametrize("shift", [0, 1])
def test_random_golden_from_reference(engine, shift):
    g, lib, q = _random_fixture()
    engine.load_library(CH, lib)
    bp, bs, npairs, pairs = engine.best_match_batch(CH, q, g["cand"].ravel(), g["cand_off"], 0.02, bool(shift))
    assert np.array_equal(bp, g[f"best_{shift}"])
    assert np.array_equal(bs, g[f"score_{shift}"])
    assert np.array_equal(npairs, g[f"npairs_{shift}"])
    for i in range(len(bp)):
        assert np.array_equal(canon_pairs(pairs[i], npairs[i]), canon_pairs(g[f"pairs_{shift}"][i], npairs[i]))


@pytest.mark.parametrize("shift", [False, True])
@pytest.mark.parametrize("tol", [0.02, 0.5])
def test_random_vs_oracle_bit_exact(engine, oracle, synth, shift, tol):
    lib = synth.make_library(4000, seed=41, decoy_seed=42)
    q = synth.make_queries(lib, 500, seed=43)
    rng = np.random.default_rng(44)
    n_lib = len(lib["prec_mz"])
    counts = rng.integers(0, 200, 500)
    counts[:3] = [0, 1, 1024]
    off = np.zeros(501, np.int64)
    np.cumsum(counts, out=off[1:])
    cand = rng.integers(0, n_lib, off[-1]).astype(np.int32)
    t = q["truth"]
    for i in range(500):
        if t[i] >= 0 and counts[i] > 0:
            cand[off[i] + rng.integers(counts[i])] = t[i]
    engine.load_library(CH, lib)
    got = engine.best_match_batch(CH, q, cand, off, tol, shift)
    want = oracle.best_match_batch(q, lib, cand, off, tol, shift, sort_mode=1)
    has = counts > 0
    assert np.array_equal(got[0], want[0])
    assert np.array_equal(got[1][has], want[1][has])
    assert np.array_equal(got[2][has], want[2][has])
    mp = min(got[3].shape[1], want[3].shape[1])
    for i in np.flatnonzero(has):
        n = got[2][i]
        assert np.array_equal(got[3][i, :n], want[3][i, :n])  # same greedy order under the total order
    assert (got[0][~has] == -1).all() and (got[2][~has] == 0).all()
    if oracle.have_ref():
        ref = oracle.ref_best_match_batch(q, lib, cand, off, tol, shift)
        assert np.array_equal(got[0][has], ref[0][has]) and np.array_equal(got[1][has], ref[1][has])
        for i in np.flatnonzero(has):
            assert np.array_equal(canon_pairs(got[3][i], got[2][i]), canon_pairs(ref[3][i], ref[2][i]))


def test_more_than_64_peaks_path(engine, oracle):
    rng = np.random.default_rng(3)
    def spec(n, prec, z):
        mz = np.sort(rng.uniform(100, 1500, n)).astype(np.float32)
        it = rng.random(n).astype(np.float32)
        return dict(prec=prec, z=z, mz=mz.tolist(), I=(it / np.linalg.norm(it)).tolist(),
                    chg=rng.integers(0, 3, n).tolist())
    cands = [spec(int(rng.integers(65, 129)), 600 + i * 0.37, 3) for i in range(40)]
    qs = [spec(100, 604.0, 3), spec(128, 610.0, 3)]
    # make query 0 share peaks with candidate 5
    qs[0]["mz"][:60] = cands[5]["mz"][:60]
    qs[0]["mz"] = sorted(qs[0]["mz"])
    q, lib = _store(qs), _store(cands)
    engine.load_library(CH, lib)
    ids = np.tile(np.arange(40, dtype=np.int32), 2)
    off = np.array([0, 40, 80], np.int64)
    got = engine.best_match_batch(CH, q, ids, off, 0.05, True)
    want = oracle.best_match_batch(q, lib, ids, off, 0.05, True, sort_mode=1, max_pairs=128)
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
    for i in range(2):
        assert np.array_equal(got[3][i, :got[2][i]], want[3][i, :want[2][i]])


@pytest.mark.parametrize("npk", [48, 110])
def test_more_tentative_matches_than_the_on_chip_list(engine, oracle, npk):
    """Dense spectra and a wide fragment tolerance: far more than 256 tentative matches per pair. The
    kernels switch to the exact greedy-by-re-enumeration path (no truncation, no error); both the
    <= 64-peak fast kernel and the general kernel are covered."""
    rng = np.random.default_rng(9)
    def spec(prec, z):
        mz = np.sort(rng.uniform(300, 330, npk)).astype(np.float32)
        it = (rng.random(npk) + 0.05).astype(np.float32)
        return dict(prec=prec, z=z, mz=mz.tolist(), I=(it / np.linalg.norm(it)).tolist(),
                    chg=rng.integers(0, 3, npk).tolist())
    cands = [spec(500 + 3.1 * i, 2) for i in range(12)]
    qs = [spec(503.0, 2), spec(520.0, 2), spec(500.0, 2)]
    q, lib = _store(qs), _store(cands)
    engine.load_library(CH, lib)
    ids = np.tile(np.arange(12, dtype=np.int32), 3)
    off = np.array([0, 12, 24, 36], np.int64)
    for shift in (False, True):
        got = engine.best_match_batch(CH, q, ids, off, 5.0, shift, max_pairs=128)
        want = oracle.best_match_batch(q, lib, ids, off, 5.0, shift, sort_mode=1, max_pairs=128)
        assert np.array_equal(got[0], want[0])
        assert np.array_equal(got[1], want[1])
        assert np.array_equal(got[2], want[2]) and got[2].min() > npk // 2
        for i in range(3):
            assert np.array_equal(got[3][i, :got[2][i]], want[3][i, :want[2][i]])


def test_capacity_errors_are_loud(engine):
    from ann_solo_b200 import SoloError
    big = dict(prec=500.0, z=2, mz=np.linspace(100, 900, 129).tolist(), I=[0.1] * 129)
    with pytest.raises(SoloError, match="128 peaks"):
        engine.load_library(CH, _store([big]))
    with pytest.raises(ValueError, match="ascending"):
        engine.load_library(CH, _store([dict(prec=500.0, z=2, mz=[200, 100], I=[.5, .5])]))
    with pytest.raises(SoloError, match="charges 0..7"):
        engine.load_library(CH, _store([dict(prec=500.0, z=9, mz=[100, 200], I=[.5, .5])]))


def test_get_best_match_api(engine, synth):
    from ann_solo_b200.spectral_library import InMemoryLibrary
    from ann_solo_b200.spectrum_match import get_best_match
    lib = synth.make_library(400, seed=51)
    q = synth.make_queries(lib, 20, seed=52)
    reader = InMemoryLibrary(lib)
    qreader = InMemoryLibrary(q)
    hits = 0
    for i in range(20):
        if q["truth"][i] < 0:
            continue
        query = qreader.read_spectrum(i)
        cands = [reader.read_spectrum(int(r)) for r in {int(q["truth"][i]), 1, 2, 3, 5, 8}]
        match, score, pm = get_best_match(query, cands, 0.02, True, engine=engine)
        assert isinstance(score, float) and all(len(p) == 2 for p in pm)
        hits += match.identifier == q["truth"][i]
        assert np.array_equal(query.charge, np.zeros(len(query.mz), np.uint8))
    assert hits >= 8
    with pytest.raises(ValueError):
        get_best_match(qreader.read_spectrum(0), [], 0.02, True, engine=engine)
