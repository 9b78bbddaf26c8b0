"""The oracle restatement pinned against fixtures minted from the reference itself
(tests/golden/make_golden.py) and, when present, against the reference's compiled core."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, canon_pairs


def test_hash_idx_matches_reference(oracle):
    g = np.load(os.path.join(GOLDEN, "vectoriser.npz"))
    got = np.array([oracle.hash_idx(b, 800) for b in range(len(g["hash_800"]))], np.uint16)
    assert np.array_equal(got, g["hash_800"])
    got400 = np.array([oracle.hash_idx(b, 400) for b in range(1000)], np.uint16)
    assert np.array_equal(got400, g["hash_400_first1000"])
    # SURVEY.md §8c probe values
    assert [oracle.hash_idx(i, 800) for i in range(10)] == [281, 427, 492, 206, 626, 541, 162, 215, 574, 344]


def test_get_dim_matches_reference(oracle):
    g = np.load(os.path.join(GOLDEN, "vectoriser.npz"))
    for mn, mx, bs, n, s, e in g["get_dim"]:
        gn, gs, ge = oracle.get_dim(mn, mx, bs)
        assert (gn, gs, ge) == (int(n), s, e)
    assert oracle.get_dim(11, 2010, 0.04) == (49976, 10.96, 2010.0)


def test_bins_follow_numpy_precision(oracle):
    g = np.load(os.path.join(GOLDEN, "vectoriser.npz"))
    mz64 = g["mz64"]
    b32 = [oracle.mz_to_bin(np.float32(m), 10.96, 0.04) for m in mz64.astype(np.float32)]
    b64 = [oracle.mz_to_bin(float(m), 10.96, 0.04) for m in mz64]
    assert np.array_equal(b32, g["bins32"])
    assert np.array_equal(b64, g["bins64"])
    assert oracle.mz_to_bin(np.float32(100.0), 10.96, 0.04) == 2226  # float32 arithmetic
    assert oracle.mz_to_bin(100.0, 10.96, 0.04) == 2225              # float64 arithmetic


@pytest.mark.parametrize("prec", ["f32", "f64"])
def test_vectorize_matches_reference(oracle, prec):
    g = np.load(os.path.join(GOLDEN, "vectoriser.npz"))
    mz = g["mz64"].astype(np.float32) if prec == "f32" else g["mz64"]
    want = g["v32"] if prec == "f32" else g["v64"]
    got = oracle.vectorize(mz, g["inten"], g["off"])
    # the occupied slots are identical; values agree to the ulp-level difference between
    # NumPy's BLAS norm and the oracle's fixed summation order: <= 2 float32 ulps, tolerance
    # 3e-7 relative (north_star allows 1e-5 on scores)
    assert np.array_equal(got != 0, want != 0)
    np.testing.assert_allclose(got, want, rtol=3e-7, atol=0)
    if prec == "f64":
        raw = oracle.vectorize(mz, g["inten"], g["off"], norm=False)
        assert np.array_equal(raw, g["v64_raw"])  # un-normalised accumulation is bit-exact


def _store(specs):
    mz, inten, chg, off, pm, pz = [], [], [], [0], [], []
    for s in specs:
        mz += list(s["mz"]); inten += list(s["I"]); chg += list(s.get("chg", [0] * len(s["mz"])))
        off.append(len(mz)); pm.append(s["prec"]); pz.append(s["z"])
    return dict(mz=np.array(mz, np.float32), inten=np.array(inten, np.float32), chg=np.array(chg, np.uint8),
                off=np.array(off, np.int64), prec_mz=np.array(pm, np.float64), prec_z=np.array(pz, np.int32))


@pytest.mark.parametrize("sort_mode", [0, 1])
def test_scorer_kats(oracle, sort_mode):
    kats = json.load(open(os.path.join(GOLDEN, "scorer_kat.json")))
    assert len(kats) >= 10
    for k in kats:
        q, lib = _store([k["query"]]), _store(k["candidates"])
        n = len(k["candidates"])
        bp, bs, npairs, pairs = oracle.best_match_batch(q, lib, np.arange(n), np.array([0, n]), k["tol"],
                                                        k["allow_shift"], sort_mode=sort_mode)
        assert bp[0] == k["best"], k["name"]
        assert bs[0] == k["score"], k["name"]
        assert pairs[0, :npairs[0]].tolist() == k["pairs"], k["name"]


def _random_fixture():
    g = np.load(os.path.join(GOLDEN, "scorer_random.npz"))
    lib = {k: g[f"lib_{k}"] for k in ("mz", "inten", "chg", "off", "prec_mz", "prec_z")}
    q = {k: g[f"q_{k}"] for k in ("mz", "inten", "chg", "off", "prec_mz", "prec_z")}
    return g, lib, q


@pytest.mark.parametrize("shift", [0, 1])
@pytest.mark.parametrize("sort_mode", [0, 1])
def test_scorer_random_golden(oracle, shift, sort_mode):
    g, lib, q = _random_fixture()
    bp, bs, npairs, pairs = oracle.best_match_batch(q, lib, g["cand"].ravel(), g["cand_off"], 0.02, bool(shift),
                                                    sort_mode=sort_mode)
    assert np.array_equal(bp, g[f"best_{shift}"])
    assert np.array_equal(bs, g[f"score_{shift}"])  # bit-exact doubles
    assert np.array_equal(npairs, g[f"npairs_{shift}"])
    for i in range(len(bp)):
        assert np.array_equal(canon_pairs(pairs[i], npairs[i]), canon_pairs(g[f"pairs_{shift}"][i], npairs[i]))


def test_scorer_port_equals_compiled_reference(oracle, synth):
    if not oracle.have_ref():
        pytest.skip("reference sources absent and oracle/_ref not prebuilt")
    lib = synth.make_library(3000, seed=5, decoy_seed=6)
    q = synth.make_queries(lib, 200, seed=7)
    rng = np.random.default_rng(8)
    n_lib = len(lib["prec_mz"])
    cand = np.sort(rng.integers(0, n_lib, (200, 64)), axis=1).astype(np.int32)
    t = q["truth"]
    cand[t >= 0, 0] = t[t >= 0]
    off = np.arange(0, 200 * 64 + 1, 64, dtype=np.int64)
    for shift in (False, True):
        r = oracle.ref_best_match_batch(q, lib, cand.ravel(), off, 0.02, shift)
        for mode in (0, 1):
            p = oracle.best_match_batch(q, lib, cand.ravel(), off, 0.02, shift, sort_mode=mode)
            assert np.array_equal(p[0], r[0]) and np.array_equal(p[1], r[1]) and np.array_equal(p[2], r[2])
            for i in range(200):
                assert np.array_equal(canon_pairs(p[3][i], p[2][i]), canon_pairs(r[3][i], r[2][i]))
    # planted candidates are found
    found = cand[np.arange(200), r[0]] == t
    assert found[t >= 0].mean() > 0.9


def test_scorer_empty_and_edge(oracle):
    q = _store([dict(prec=500.0, z=2, mz=[100, 200, 300, 400], I=[.5] * 4)])
    lib = _store([dict(prec=490, z=2, mz=[100, 200, 290, 380], I=[.5] * 4)])
    bp, bs, npairs, _ = oracle.best_match_batch(q, lib, np.zeros(0, np.int32), np.array([0, 0]), 0.02, True)
    assert bp[0] == -1 and npairs[0] == -1
